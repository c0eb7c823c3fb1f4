// la3dm_b200 -- GPOctoMap::predict on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// GPRegressor::predict (include/gpoctomap/gpregressor.h:80-92) for one trained data block and every test leaf that has
// it in its ExtendedBlock:   m = Ks^T alpha,   v = L^-1 Ks,   var = sf2 - diag(v^T v).
// v is a triangular solve with many right-hand sides (one per test leaf) -- a blocked TRSM.  With X = V^T (leaves x n):
//     X[:, bi] = ( Ks^T[:, bi] - sum_{bj < bi} X[:, bj] L[bi, bj]^T ) L[bi, bi]^-T        (blocks of 16 points)
// the off-diagonal update is a GEMM and runs on tcgen05.mma (accumulators in TMEM); the 16 x 16 diagonal solve is a
// forward substitution per leaf on the SIMT lanes, like the scalar kernel's.  (Forming L^-1 and doing ONE GEMM was
// tried first: |L^-1| ~ 10 amplifies the rounding of the products into var = sf2 - |v|^2 ~ 1e-3 and moved m / var by
// 3e-3 -- the substitution does not amplify, which is why upstream's LLT solve works in fp32 at all.)
//   * one CTA walks data blocks; per block L goes to shared memory once (split for the tensor cores), then tiles of 128
//     test leaves (two test blocks of 64 finest slots): thread r computes row r of Ks^T and the mean m = Ks^T alpha in
//     the scalar path's order (bit-identical to it), then for every block of 16 points: one elected thread issues the
//     MMAs D = X[:, < i0] L[i0 .. i0 + 16, < i0]^T, thread r reads TMEM lane r back with tcgen05.ld, finishes its 16
//     unknowns and appends them (split) to the A operand for the next block; |v|^2 accumulates in ascending i;
//   * fp32 through TF32 tensor cores: every operand is split exactly into three TF32 terms (11 + 11 + 2 significant
//     bits: hi = x with 13 mantissa bits cleared, mid likewise of x - hi, lo = the rest) and the six products whose
//     weight reaches 2^-24 are accumulated (hi hi, hi mid, mid hi, mid mid, hi lo, lo hi): products as exact as fp32's,
//     accumulation in fp32 inside the tensor core -- plain TF32 (2^-11) would be useless against cond(K) ~ 1e4.  The
//     order of additions differs from the scalar path's like the reference's own R-tree order does; the budget is the
//     measured distribution in profiles/r2_parity_gp_vs_ref.json;
//   * operands are K-major, no swizzle: 8 x 16-byte core matrices, LBO = 128 B between the two halves of a K = 8 step,
//     SBO = K_pad * 32 B between groups of 8 rows (cute::UMMA canonical INTERLEAVE layout).
// Data blocks of more than kTcMaxN points (rare: a block holds at most 64 hit + 64 free cells) stay on the SIMT kernel.
#include "block_common.cuh"
#include "gp_common.cuh"

namespace la3dm_b200 {

namespace {

constexpr int kTcMaxN = 64;                      // points per data block handled here
constexpr int kTcKPad = 64;                      // K of the operand buffers (multiple of 8)
constexpr int kTcBlk = 16;                       // points per TRSM block = N of an MMA
constexpr int kTcM = 128;                        // leaves per tile = TMEM lanes
constexpr int kTcCols = 32;                      // TMEM columns allocated (power of two >= 32 >= kTcBlk)

struct TcSmem {
    float A[3][kTcM * kTcKPad];                  // Ks^T, split hi / mid / lo, canonical K-major layout
    float B[3][kTcMaxN * kTcKPad];               // L (row i, column k), split, same layout
    float Ks[kTcM][kTcMaxN + 1];                 // Ks^T of the tile in fp32 (padded rows: no bank conflicts)
    float L[kTcMaxN * (kTcMaxN + 1) / 2 + kTcMaxN];   // packed L, then alpha (k_gp_train's storage)
    float4 x[kTcMaxN];                           // the block's training points (pre-scaled)
    unsigned long long bar;                      // mbarrier the MMAs commit to
    unsigned int tmem;                           // TMEM base address written by tcgen05.alloc
    int unit_t[8], unit_nb[8], n_units;
};

// byte offset of element (row, k) in a K-major no-swizzle operand with K = kTcKPad
__device__ __forceinline__ int canon(int row, int k) {
    return ((row >> 3) * (kTcKPad * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3));   // in floats
}

__device__ __forceinline__ void split3(float v, float &hi, float &mid, float &lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float r = v - hi;                      // exact
    mid = __uint_as_float(__float_as_uint(r) & 0xFFFFE000u);
    lo = r - mid;                                // exact, at most 2 significant bits
}

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int) __cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ unsigned long long make_desc(unsigned int saddr) {
    const unsigned long long lbo = 128 >> 4, sbo = (kTcKPad * 32) >> 4;
    return (unsigned long long) ((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(unsigned int tmem_d, unsigned long long da, unsigned long long db, unsigned int idesc,
                                         unsigned int accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(kTcM, 1)
k_gp_mv_tc(const NeighbourPlan *__restrict__ plan, const float4 *__restrict__ pts,
           const unsigned int *__restrict__ db_id, const unsigned int *__restrict__ db_start,
           const unsigned int *__restrict__ cell_test, const GridDesc *__restrict__ g,
           const unsigned long long *__restrict__ off, const float *__restrict__ store,
           const long long *__restrict__ keys, const unsigned char *__restrict__ pool, const float3 *__restrict__ lut,
           const DevParams *__restrict__ Pg, const ScanArgs *__restrict__ A, const ScanCounters *__restrict__ cnt,
           unsigned int t0, unsigned int chunk, float2 *mv) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    TcSmem &S = *reinterpret_cast<TcSmem *>(tc_smem_raw);
    __shared__ DevParams Ps;
    load_params(Ps, Pg);
    const DevParams &P = Ps;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool dead = cnt->overflow != 0 || t0 >= cnt->n_test_blocks;
    // ---- TMEM columns for the accumulator tile, the mbarrier
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&S.tmem)), "n"(kTcCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned int tmem = S.tmem;
    unsigned int phase = 0;

    const unsigned int T = cnt->n_test_blocks, t1 = min(T, t0 + chunk);
    const unsigned int D = dead ? 0u : cnt->n_data_blocks;
    const float sf2 = P.sf2, bs = P.block_size;
    const float scale = (float) (1.73205 / (double) P.ell);    // gpregressor.h:115
    const unsigned int shard_world = (unsigned int) A->shard_world, shard_rank = (unsigned int) A->shard_rank;
    const int pruned = P.pruned_state;
    const int nz = g->n[2], ny = g->n[1], nx = g->n[0];
    const int finest = P.finest;                              // 64 at block_depth 3 (this kernel: finest <= 64)
    const int ddx[7] = {0, 1, -1, 0, 0, 0, 0}, ddy[7] = {0, 0, 0, 1, -1, 0, 0}, ddz[7] = {0, 0, 0, 0, 0, 1, -1};

    for (unsigned int d = blockIdx.x; d < D; d += gridDim.x) {
        const unsigned int first = db_start[d], n = db_start[d + 1] - first;
        if (n == 0 || n > (unsigned int) kTcMaxN) continue;
        // ---- the test blocks that have this data block as neighbour nb: the cell  own - dir[nb]
        if (tid == 0) {
            const unsigned int id = db_id[d];
            const int z = (int) (id % (unsigned int) nz), y = (int) ((id / (unsigned int) nz) % (unsigned int) ny),
                      x = (int) (id / ((unsigned int) nz * (unsigned int) ny));
            int nu = 0;
            for (int nb = 0; nb < 7; ++nb) {
                const int xx = x - ddx[nb], yy = y - ddy[nb], zz = z - ddz[nb];
                if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                const unsigned int c = ((unsigned int) xx * (unsigned int) ny + (unsigned int) yy) * (unsigned int) nz + (unsigned int) zz;
                const unsigned int tt = cell_test[c];
                if (tt == 0) continue;
                const unsigned int t = tt - 1;
                if (t < t0 || t >= t1 || t % shard_world != shard_rank) continue;
                S.unit_t[nu] = (int) t; S.unit_nb[nu] = nb; ++nu;
            }
            S.n_units = nu;
        }
        __syncthreads();
        const int n_units = S.n_units;
        if (n_units == 0) { __syncthreads(); continue; }
        // ---- L, alpha, the points
        {
            const float *Lg = store + off[d];
            const unsigned int words = n * (n + 1) / 2 + n;
            for (unsigned int w = tid; w < words; w += kTcM) S.L[w] = Lg[w];
            if ((unsigned int) tid < n) S.x[tid] = pts[first + tid];
        }
        __syncthreads();
        // ---- B = L (zero above the diagonal), split, canonical layout
        for (int e = tid; e < kTcMaxN * kTcKPad; e += kTcM) {
            const int row = e / kTcKPad, k = e - row * kTcKPad;
            const float v = ((unsigned int) row < n && k <= row) ? S.L[row * (row + 1) / 2 + k] : 0.f;
            float hi, mid, lo;
            split3(v, hi, mid, lo);
            const int o = canon(row, k);
            S.B[0][o] = hi; S.B[1][o] = mid; S.B[2][o] = lo;
        }
        const float *alpha = S.L + n * (n + 1) / 2;
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
        const unsigned int idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int) (kTcBlk >> 3) << 17) | ((unsigned int) (kTcM >> 4) << 24);

        for (int u0 = 0; u0 < n_units; u0 += 2) {
            // ---- thread r = leaf slot r & 63 of unit u0 + (r >> 6): its row of Ks^T and the mean (scalar path's order)
            const int u = u0 + (tid >> 6), j = tid & 63;
            int node = -1, t = -1, nb = 0;
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (u < n_units && j < finest) {
                t = S.unit_t[u]; nb = S.unit_nb[u];
                const NeighbourPlan *pl = plan + t;
                const unsigned int slot = pl->slot;
                if (pl->is_new) node = P.layer_off[P.depth - 1] + j;
                else {
                    const unsigned char *rst = pool + (size_t) slot * (size_t) P.rec_bytes + P.st_off;
                    int dd = P.depth - 1, i = j, shift = 0;
                    while (dd > 0 && (rst[P.layer_off[dd] + i] & 7) == pruned) { --dd; i >>= 3; shift += 3; }
                    if (((i << shift) == j) && ((rst[P.layer_off[dd] + i] & 7) != pruned)) node = P.layer_off[dd] + i;
                }
                if (node >= 0) {
                    const long long key = keys[slot];
                    const float cx = axis_center(key >> 40, bs), cy = axis_center((key >> 20) & 0xFFFFF, bs),
                                cz = axis_center(key & 0xFFFFF, bs);
                    const float3 o = lut[node];
                    qx = scale * (o.x + cx); qy = scale * (o.y + cy); qz = scale * (o.z + cz);
                }
            }
            float mu = 0.f, v2 = 0.f;
            for (unsigned int k = 0; k < n; ++k) {
                float v = 0.f;
                if (node >= 0) {
                    const float4 xk = S.x[k];
                    v = matern3(xk.x, xk.y, xk.z, qx, qy, qz, sf2);
                    mu += v * alpha[k];
                }
                S.Ks[tid][k] = v;
            }
            // ---- blocked forward substitution, 16 unknowns per step
            for (int i0 = 0; i0 < (int) n; i0 += kTcBlk) {
                float rhs[kTcBlk];
#pragma unroll
                for (int q = 0; q < kTcBlk; ++q) rhs[q] = (unsigned int) (i0 + q) < n ? S.Ks[tid][i0 + q] : 0.f;
                if (i0 > 0) {
                    // D[leaf][q] = sum_{k < i0} X[leaf][k] L[i0 + q][k] on the tensor cores
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // A / B stores -> async proxy
                    __syncthreads();
                    if (tid == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;\n");
                        const int ai[6] = {0, 0, 1, 1, 0, 2}, bi[6] = {0, 1, 0, 1, 2, 0};
                        unsigned int acc = 0;
                        const unsigned int b_row = (unsigned int) (i0 >> 3) * (unsigned int) (kTcKPad * 32);   // SBO per 8 rows
                        for (int ks = 0; ks < (i0 >> 3); ++ks)
                            for (int q = 0; q < 6; ++q) {
                                const unsigned long long da = make_desc(smem_u32(S.A[ai[q]]) + (unsigned int) ks * 256u);
                                const unsigned long long db = make_desc(smem_u32(S.B[bi[q]]) + b_row + (unsigned int) ks * 256u);
                                mma_tf32(tmem, da, db, idesc, acc);
                                acc = 1;
                            }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&S.bar)) : "memory");
                    }
                    {
                        unsigned int done = 0, spins = 0;
                        while (!done) {
                            if (++spins > (1u << 24)) __trap();          // (an MMA that never completes must not hang the device)
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                "selp.u32 %0, 1, 0, p;\n\t}\n"
                                : "=r"(done) : "r"(smem_u32(&S.bar)), "r"(phase) : "memory");
                        }
                        phase ^= 1u;
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;\n");
                    // thread r owns TMEM lane r (warp w may touch lanes 32 w .. 32 w + 31)
                    unsigned int r[16];
                    const unsigned int taddr = tmem + ((unsigned int) (warp * 32) << 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int q = 0; q < kTcBlk; ++q) rhs[q] -= __uint_as_float(r[q]);
                    asm volatile("tcgen05.fence::before_thread_sync;\n");
                }
                // diagonal block: x_q = (rhs_q - sum_{p < q} L[i0 + q][i0 + p] x_p) / L[i0 + q][i0 + q]
#pragma unroll
                for (int q = 0; q < kTcBlk; ++q) {
                    const int i = i0 + q;
                    float xq = 0.f;
                    if ((unsigned int) i < n) {
                        const float *ri = S.L + i * (i + 1) / 2 + i0;
                        float sacc = rhs[q];
#pragma unroll
                        for (int pq = 0; pq < q; ++pq) sacc -= ri[pq] * rhs[pq];
                        xq = sacc / ri[q];
                        v2 += xq * xq;
                    }
                    rhs[q] = xq;                                       // (rhs[0 .. q] now holds the solved unknowns)
                    float hi, mid, lo;
                    split3(xq, hi, mid, lo);
                    const int o = canon(tid, i);
                    S.A[0][o] = hi; S.A[1][o] = mid; S.A[2][o] = lo;
                }
            }
            if (node >= 0) {
                const int groups32 = ((finest + 31) / 32) * 32;
                mv[((size_t) ((unsigned int) t - t0) * 7 + (unsigned int) nb) * (size_t) groups32 + (unsigned int) j] =
                    make_float2(mu, sf2 - v2);
            }
            // the next tile overwrites Ks, A and the accumulators
            __syncthreads();
        }
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTcCols));
}

}  // namespace

size_t gp_tc_smem_bytes() { return sizeof(TcSmem) + 128; }
int gp_tc_max_n() { return kTcMaxN; }

// (mean, variance) of the units whose neighbour regressor has at most kTcMaxN points, on the tensor cores
void Map::enqueue_gp_mv_tc(unsigned int t0, unsigned int chunk) {
    LA3DM_CUDA(cudaFuncSetAttribute(k_gp_mv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gp_tc_smem_bytes()));
    k_gp_mv_tc<<<num_sms, kTcM, gp_tc_smem_bytes(), stream>>>(
        plan.as<NeighbourPlan>(), pts_sorted.as<float4>(), db_id.as<unsigned int>(), db_start.as<unsigned int>(),
        cell_test.as<unsigned int>(), d_grid, gp_off.as<unsigned long long>(), gp_store.as<float>(), keys.as<long long>(),
        pool.as<unsigned char>(), d_lut, d_params, d_args, d_cnt, t0, chunk, gp_mv.as<float2>());
    ++launches;
}

}  // namespace la3dm_b200
