// Shared declarations of the MSM translation units (msm.cu: Pippenger pipeline; msm_tree.cu: batched-
// affine pairwise rounds).
#pragma once
#include "common.cuh"

namespace eon {

constexpr int MSM_THREADS = 256;
constexpr u32 SIGN_BIT = 0x80000000u;

struct MsmShape {
  u32 c;        // window bits (2..20)
  u32 W;        // windows = ceil(256 / c)
  u32 NB;       // buckets per bucket set = 2^(c-1)
  u32 nsets;    // bucket sets per column: W (plain) or 1 (merged: windows share one set because
                // window t reads the precomputed table 2^(c t) * P_i instead of P_i)
  u32 merged;
  u32 chunk;    // buckets per reduction chunk
  u32 nchunks;  // NB / chunk
  u64 tab_stride;  // merged: points per table level
  u64 base_first;  // merged: index of this MSM's point 0 inside a table level
  u64 seg_cap;     // entry slots per segment: n (plain) or n * W (merged), plus alignment padding
  u32 rounds;      // batched-affine pairwise rounds before the serial XYZZ finisher (0 = none);
                   // bucket starts are aligned to 2^rounds entry slots
};

constexpr u32 ENTRY_NONE = 0xffffffffu;  // unused entry slot (reads as the identity point)

// R rounds of out[j] = in[2j] + in[2j+1] over the flat slot array (all segments), affine with one
// shared inversion per round.  Round 0 reads entries/bases, later rounds read points.  On return
// *out_pts points to the final array of (total_slots >> rounds) affine points.
int msm_tree_rounds(eon_ctx* ctx, const G1Affine* d_bases, const u32* d_entries, u64 total_slots, u32 rounds,
                    const G1Affine** out_pts);



}  // namespace eon
