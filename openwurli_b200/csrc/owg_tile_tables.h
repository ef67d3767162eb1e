// Term tables of the lane-tiled DK step (owg_tile.cuh): build_rhs (gen_preamp.rs:3041-3095) row by row.
//
// A render instance is owned by a tile of 4 lanes; lane q owns three rows of the 12-node system, one per "position":
//
//        position 0   position 1   position 2
//   q=0      3            1            0
//   q=1      2            5            9
//   q=2      4            8           10
//   q=3      7            6           11
//
// The assignment does three things at once.  (1) Rows of similar length share a position (7/7/5/5, 4/4/3/3 and 3/2/2/1 terms), so
// the branch-free instruction stream needs 7 + 4 + 3 = 14 term slots per lane instead of 18.  (2) The Newton solve is split by
// DEVICE ROW over lanes 0..2 (lane 3 mirrors lane 0), and the junction voltages p = N_v * v_pred (gen_preamp.rs:3111-3120) are
// p0 = -v_pred[2], p1 = v_pred[2] - v_pred[5], p2 = v_pred[4] - v_pred[8]: lanes 1 and 2 find theirs as (position 0) - (position 1)
// of their own rows, lane 0 fetches v_pred[2] with one shuffle.  (3) The output node (row 10) and the V-source row (row 11, which
// the ringing / damping tests skip) sit in position 2.
//
// Every row of build_rhs is a left-to-right sum of (coefficient x value) products: the structural non-zeros of a_neg times v_prev in
// column order, then the N_i * i_nl_prev terms in device order (the reference adds them with separate `+=`).  Rows are
// padded to the slot count of their position with the product (-0.0 x 1.0) = -0.0, whose addition is the identity on every f64
// (including +-0, infinities and NaN), so all lanes run one branch-free instruction stream and still produce the reference's bits.
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
// value sources (ROW order; the kernel's gather buffer is position-major, see OWG_TILE_LOC): 0..11 = v_prev, 12..14 = i_nl_prev,
// 15 = the constant 1.0
#define OWG_TX_IL 12
#define OWG_TX_ONE 15
#define OWG_TX_PP 16  // i_nl_prev_prev in the kernel's home buffer (not a build_rhs operand)

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

// row owned by lane q at position p, and the inverse: where row r lives in the position-major gather buffer (p * 4 + q)
#define OWG_TILE_ROWS_INIT {{3, 1, 0}, {2, 5, 9}, {4, 8, 10}, {7, 6, 11}}
#define OWG_TILE_LOC_INIT {8, 4, 1, 0, 2, 5, 7, 3, 6, 9, 10, 11, 12, 13, 14, 15}

// Slots per lane: position 0 uses all 7 terms, position 1 at most 4 (slots 7..10), position 2 at most 3 (slots 11..13).
#define OWG_TILE_SLOTS 14
#define OWG_TILE_SLOT_POS(s) ((s) < 7 ? 0 : ((s) < 11 ? 1 : 2))
#define OWG_TILE_SLOT_K(s) ((s) < 7 ? (s) : ((s) < 11 ? (s) - 7 : (s) - 11))
// a_neg[6][6] (the only R_ldr-dependent entry): row 6 = lane 3, position 1, term 1
#define OWG_TILE_AN66_LANE 3
#define OWG_TILE_AN66_SLOT 8
