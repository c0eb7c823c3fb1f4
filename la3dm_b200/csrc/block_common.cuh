// la3dm_b200 -- pieces shared by the per-test-block predict kernels (one warp per block, block_depth <= 3):
// staging of the block record in shared memory, leaf resolution, OcTree::prune and the write-back.
#pragma once
#include "engine.cuh"

namespace la3dm_b200 {

constexpr int kRecMax = 672;          // bytes of a depth-3 record: 73 * 8 + 73 -> 16-byte multiple

// sinf(y) and cosf(y) for y in [0, 120), BIT FOR BIT what the host libm returns: glibc >= 2.28 computes them in
// double precision (ARM optimized-routines sincosf: quadrant n = round(y * 2 / pi) from a 2^24-scaled product,
// x = y - n * (pi / 2), a degree-7 / degree-8 polynomial pair in double, one rounding to float).  The constants below
// are the published ones; tests/test_oracle_golden.py::test_sincos_restatement_matches_libm checks a CPU copy of this
// function against libm on 4e5 arguments of the form  d * 2 * 3.1415926f.  Every product and sum is rounded
// separately (__dmul_rn / __dadd_rn are never contracted), like the x86-64 baseline build.
// Why it matters: a voxel seen by one free point flips UNKNOWN -> FREE at k = 1.33e-3 (p = 0.001 / (0.002 + k) < 0.3),
// a flipped state can complete a group of 8 equal siblings, and a prune changes the leaf SET -- which must stay
// bit-exact.  With sin / cos identical to the reference's, only the order of the per-leaf additions differs.
__device__ __forceinline__ void sincosf_libm(float y, float &s, float &c) {
    const double x0 = (double) y;
    const int n = (__double2int_rz(__dmul_rn(x0, 0x1.45F306DC9C883p+23)) + 0x800000) >> 24;
    double x = __dadd_rn(x0, -__dmul_rn((double) n, 0x1.921FB54442D18p0));
    const double x2 = __dmul_rn(x, x);
    if (((n + 1) >> 1) & 1) x = -x;                          // sign[n & 3] = {1, -1, -1, 1}
    // sine polynomial
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __dadd_rn(0x1.1107605230bc4p-7, __dmul_rn(x2, -0x1.994eb3774cf24p-13));
    const double x7 = __dmul_rn(x3, x2);
    const double sa = __dadd_rn(x, __dmul_rn(x3, -0x1.555545995a603p-3));
    const double sp = __dadd_rn(sa, __dmul_rn(x7, s1));
    // cosine polynomial (the table of the quadrants with n & 2 holds the negated coefficients: the result is negated)
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(-0x1.6c087e89a359dp-10, __dmul_rn(x2, 0x1.99343027bf8c3p-16));
    const double c1 = __dadd_rn(1.0, __dmul_rn(x2, -0x1.ffffffd0c621cp-2));
    const double x6 = __dmul_rn(x4, x2);
    const double ca = __dadd_rn(c1, __dmul_rn(x4, 0x1.55553e1068f19p-5));
    double cp = __dadd_rn(ca, __dmul_rn(x6, c2));
    if (n & 2) cp = -cp;
    const float sf = (float) sp, cf = (float) cp;
    s = (n & 1) ? cf : sf;                                   // sinf: sinf_poly(x, x2, p, n)
    c = (n & 1) ? sf : cf;                                   // cosf: sinf_poly(x, x2, p, n ^ 1)
}

// default record in shared memory: every node = (prior_A, prior_B | 0, min_ivar), UNKNOWN, !classified
// (bgkoctree_node.h:34 / gpoctree_node.h:34); the spare bytes behind the states are zero except the leaf count
__device__ __forceinline__ void stage_default_record(uint4 *srec, const DevParams &P, int lane) {
    float2 *rab = reinterpret_cast<float2 *>(srec);
    unsigned char *rb = reinterpret_cast<unsigned char *>(srec);
    for (int n = lane; n < P.nodes; n += 32) { rab[n] = make_float2(P.def_a, P.def_b); rb[P.st_off + n] = LA3DM_UNKNOWN; }
    for (int n = P.st_off + P.nodes + lane; n < P.rec_bytes; n += 32) rb[n] = n == P.st_off + P.nodes ? (unsigned char) P.finest : 0;
}

// record -> shared memory (a fresh Block gets the default node everywhere: bgkoctree_node.h:34 / gpoctree_node.h:34)
__device__ __forceinline__ void stage_record(uint4 *srec, const uint4 *grec, bool is_new, const DevParams &P, int lane) {
    float2 *rab = reinterpret_cast<float2 *>(srec);
    unsigned char *rb = reinterpret_cast<unsigned char *>(srec);
    if (is_new) {
        for (int n = lane; n < P.nodes; n += 32) { rab[n] = make_float2(P.def_a, P.def_b); rb[P.st_off + n] = LA3DM_UNKNOWN; }
        for (int n = P.st_off + P.nodes + lane; n < P.rec_bytes; n += 32) rb[n] = 0;
    } else {
        for (int w = lane; w < (P.rec_bytes >> 4); w += 32) srec[w] = grec[w];
    }
}

// Leaves of the block owned by this lane: finest slots lane and lane + 32; a slot whose ancestors were pruned resolves
// to the coarser leaf (d, i), owned by the lane of its first finest descendant (is_leaf: bgkoctree.cpp:72-82).
__device__ __forceinline__ void resolve_leaves(const unsigned char *rst, const DevParams &P, int lane, int node[2]) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int j = lane + 32 * s;
        node[s] = -1;
        if (j < P.finest) {
            int d = P.depth - 1, i = j, shift = 0;
            while (d > 0 && (rst[P.layer_off[d] + i] & 7) == kStPRUNED) { --d; i >>= 3; shift += 3; }
            if (((i << shift) == j) && ((rst[P.layer_off[d] + i] & 7) != kStPRUNED)) node[s] = P.layer_off[d] + i;
        }
    }
}

// OcTree::prune (bgkoctree.cpp:101-148) on the staged record: deepest layer first; 8 equal FREE/OCCUPIED siblings
// collapse into the parent (copy of child 0's two floats and state -- `classified` is not copied,
// bgkoctree_node.h:40-45); ends with a __syncwarp
__device__ __forceinline__ void prune_record(float2 *rab, unsigned char *rst, const DevParams &P, int lane) {
    for (int d = P.depth - 1; d > 0; --d) {
        const int off = P.layer_off[d], poff = P.layer_off[d - 1];
        const int groups = 1 << (3 * (d - 1));
        for (int g = lane; g < groups; g += 32) {
            const unsigned char s0 = rst[off + 8 * g] & 7;
            if (s0 == LA3DM_FREE || s0 == LA3DM_OCCUPIED) {
                bool same = true;
#pragma unroll
                for (int i = 1; i < 8; ++i) same = same && ((rst[off + 8 * g + i] & 7) == s0);
                if (same) {
                    rab[poff + g] = rab[off + 8 * g];
                    rst[poff + g] = (rst[poff + g] & 0x80) | s0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) rst[off + 8 * g + i] = (rst[off + 8 * g + i] & 0x80) | kStPRUNED;
                }
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void load_params(DevParams &Ps, const DevParams *Pg) {
    if (threadIdx.x < sizeof(DevParams) / 4)
        reinterpret_cast<int *>(&Ps)[threadIdx.x] = reinterpret_cast<const int *>(Pg)[threadIdx.x];
    __syncthreads();
}

}  // namespace la3dm_b200
