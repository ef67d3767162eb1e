"""Command-line mirrors of the reference's offline render tools (same flags, batched on the GPU)."""
