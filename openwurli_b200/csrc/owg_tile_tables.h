// Term tables of the lane-tiled DK step (owg_tile.cuh): build_rhs (gen_preamp.rs:3041-3095) row by row.
//
// A render instance is owned by a tile of 4 lanes; lane q owns rows q, q+4, q+8 of the 12-node system.  Every row of
// build_rhs is a left-to-right sum of (coefficient x value) products: the structural non-zeros of a_neg times v_prev in
// column order, then the N_i * i_nl_prev terms in device order (the reference adds them with separate `+=`).  Rows are
// padded to a common length with the product (-0.0 x 1.0) = -0.0, whose addition is the identity on every f64 (including
// +-0, infinities and NaN), so all lanes run one branch-free instruction stream and still produce the reference's bits.
//
// Plain C: included by the CUDA kernel and by tests/tile_tables_check.cpp (host check against the straightforward rows).
#pragma once

// coefficient sources: 0..37 = index into the group's sparse a_neg list (OWG_AN_SPARSE order; 24 = [6][6], which the kernel
// replaces by the record's per-sample an66), then constants
#define OWG_TC_NI02 38
#define OWG_TC_NI12 39
#define OWG_TC_NI14 40
#define OWG_TC_NI24 41
#define OWG_TC_NI15 42
#define OWG_TC_NI27 43
#define OWG_TC_NI28 44
#define OWG_TC_RHS11 45
#define OWG_TC_PAD 46
#define OWG_TC_COUNT 47
// value sources: 0..11 = v_prev, 12..14 = i_nl_prev, 15 = the constant 1.0
#define OWG_TX_IL 12
#define OWG_TX_ONE 15

#define OWG_TILE_ROW_TERMS 7
struct OwgRhsTerm { unsigned char c, x; };

#define OWG_T_PAD {OWG_TC_PAD, OWG_TX_ONE}
#define OWG_RHS_ROWS_INIT {                                                                                                        \
    /* 0*/ {{0, 0}, {1, 1}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                                 \
    /* 1*/ {{2, 0}, {3, 1}, {4, 2}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                                    \
    /* 2*/ {{5, 1}, {6, 2}, {7, 3}, {8, 4}, {9, 5}, {OWG_TC_NI02, 12}, {OWG_TC_NI12, 13}},                                          \
    /* 3*/ {{10, 2}, {11, 3}, {12, 4}, {13, 7}, {14, 11}, OWG_T_PAD, OWG_T_PAD},                                                    \
    /* 4*/ {{15, 2}, {16, 3}, {17, 4}, {18, 7}, {19, 8}, {OWG_TC_NI14, 13}, {OWG_TC_NI24, 14}},                                     \
    /* 5*/ {{20, 2}, {21, 5}, {22, 6}, {OWG_TC_NI15, 13}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                         \
    /* 6*/ {{23, 5}, {24, 6}, {25, 10}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                                \
    /* 7*/ {{26, 3}, {27, 4}, {28, 7}, {29, 10}, {OWG_TC_NI27, 14}, OWG_T_PAD, OWG_T_PAD},                                          \
    /* 8*/ {{30, 4}, {31, 8}, {32, 9}, {OWG_TC_NI28, 14}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                         \
    /* 9*/ {{33, 8}, {34, 9}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                               \
    /*10*/ {{35, 6}, {36, 7}, {37, 10}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD},                                                \
    /*11*/ {{OWG_TC_RHS11, OWG_TX_ONE}, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD, OWG_T_PAD}}

// Slots per lane: rows q and q+4 use all 7 terms, rows 8..11 have at most 4 (slots 14..17).
#define OWG_TILE_SLOTS 18
#define OWG_TILE_SLOT_ROW(q, s) ((q) + 4 * ((s) < 7 ? 0 : ((s) < 14 ? 1 : 2)))
#define OWG_TILE_SLOT_K(s) ((s) < 7 ? (s) : ((s) < 14 ? (s) - 7 : (s) - 14))
