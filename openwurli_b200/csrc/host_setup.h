// Host-side note-on parameterisation (see host_setup.cpp).
#pragma once
#include "../../include/owgpu.h"
#include "owg_records.h"

namespace owg {
// Voice::note_on(+overrides) -> flat init record (voice.rs:28-142).
void make_voice_init(const owg_voice_job& job, OwgVoiceInit* out);
// Speaker / volume parameters of one `preamp-bench render` job (main.rs:478-496).
void make_chain_init(const owg_bench_job& job, int group, OwgChainInit* out);
// Attack-noise raised-cosine fade-in table (hammer.rs:161-168), 16 entries.
void noise_fade_table(double* t16);
}  // namespace owg
