// la3dm_b200 -- map lifetime, the per-scan driver and the read side (export of blocks / leaves).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "engine.cuh"
#include "leaf.cuh"

namespace la3dm_b200 {

namespace {
constexpr int kThreads = 256;

// state numbering differs for BGKLV only in the PRUNED slot; bits: 0..2 state, 7 classified
__global__ void k_pack_nodes(const unsigned char *__restrict__ pool, const unsigned int *__restrict__ order,
                             unsigned int n_blocks, int nodes, int st_off, int rec_bytes, la3dm_node *out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) n_blocks * nodes) return;
    const unsigned int b = (unsigned int) (i / nodes), n = (unsigned int) (i % nodes);
    const unsigned char *rec = pool + (size_t) order[b] * rec_bytes;
    const float2 v = reinterpret_cast<const float2 *>(rec)[n];
    const unsigned char s = rec[st_off + n];
    la3dm_node o;
    o.classified = s >> 7; o._pad0[0] = o._pad0[1] = o._pad0[2] = 0;
    o.a = v.x; o.b = v.y;
    o.state = s & 7; o._pad1[0] = o._pad1[1] = o._pad1[2] = 0;
    out[i] = o;
}

__global__ void k_iota(unsigned int *v, unsigned int n) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void k_gather_keys(const long long *__restrict__ keys, const unsigned int *__restrict__ order,
                              unsigned int n, long long *out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keys[order[i]];
}

// state_mask: bit s set = leaves in state s are wanted (0xFF: all)
__device__ __forceinline__ bool leaf_wanted(const unsigned char *bst, const DevParams &P, int d, int i, unsigned int state_mask) {
    return node_is_leaf(bst, P, d, i) && ((state_mask >> (bst[P.layer_off[d] + i] & 7)) & 1u);
}

__global__ void k_leaf_count(const unsigned char *__restrict__ pool, const unsigned int *__restrict__ order,
                             unsigned int n_blocks, const DevParams *__restrict__ Pg, unsigned int state_mask,
                             unsigned int *cnt) {
    // one warp per block
    const unsigned int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_blocks) return;
    const DevParams &P = *Pg;
    const unsigned char *bst = pool + (size_t) order[w] * P.rec_bytes + P.st_off;
    unsigned int c = 0;
    for (int d = 0; d < P.depth; ++d) {
        const int n = 1 << (3 * d);
        for (int i = lane; i < n; i += 32) c += leaf_wanted(bst, P, d, i, state_mask) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[w] = c;
}

// leaves of one block in (depth, index) order
__global__ void k_leaf_fill(const unsigned char *__restrict__ pool, const long long *__restrict__ keys,
                            const unsigned int *__restrict__ order, unsigned int n_blocks,
                            const DevParams *__restrict__ Pg, const float3 *__restrict__ lut,
                            const unsigned int *__restrict__ off, unsigned int state_mask, la3dm_leaf *out) {
    const unsigned int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_blocks) return;
    const DevParams &P = *Pg;
    const unsigned int slot = order[w];
    const unsigned char *rec = pool + (size_t) slot * P.rec_bytes;
    const unsigned char *bst = rec + P.st_off;
    const float2 *bab = reinterpret_cast<const float2 *>(rec);
    const long long key = keys[slot];
    const float cx = axis_center(key >> 40, P.block_size), cy = axis_center((key >> 20) & 0xFFFFF, P.block_size),
                cz = axis_center(key & 0xFFFFF, P.block_size);
    unsigned int base = off[w];
    for (int d = 0; d < P.depth; ++d) {
        const int n = 1 << (3 * d);
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const bool leaf = i < n && leaf_wanted(bst, P, d, i, state_mask);
            const unsigned int m = __ballot_sync(0xffffffffu, leaf);
            if (leaf) {
                const unsigned int pos = base + __popc(m & ((1u << lane) - 1u));
                const int node = P.layer_off[d] + i;
                const float2 v = bab[node];
                const unsigned char s = bst[node];
                const la3dm_leaf L = make_leaf(P, key, d, i, v, s, lut[node], cx, cy, cz);
                out[pos] = L;
            }
            base += __popc(m);
        }
    }
}

__global__ void k_leaf_total(const unsigned int *__restrict__ cnt, const unsigned int *__restrict__ off,
                             unsigned int n, ScanCounters *c) {
    c->n_leaves = n ? cnt[n - 1] + off[n - 1] : 0;
}

// slots of the blocks touched since the last clearing export, with their keys (unordered)
__global__ void k_touched_collect(const unsigned char *__restrict__ touched, const long long *__restrict__ keys,
                                  unsigned int n_blocks, long long *out_keys, unsigned int *out_slots, unsigned int *count) {
    const unsigned int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool t = slot < n_blocks && touched[slot] != 0;
    const unsigned int m = __ballot_sync(0xffffffffu, t);
    unsigned int base = 0;
    if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(count, (unsigned int) __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (t) {
        const unsigned int pos = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        out_keys[pos] = keys[slot];
        out_slots[pos] = slot;
    }
}

unsigned int grow_to(unsigned int need, unsigned int floor_) {
    const unsigned long long w = (unsigned long long) need + need / 8 + 64;
    return (unsigned int) std::min<unsigned long long>(std::max<unsigned long long>(w, floor_), 0x7FFFFFF0ull);
}

}  // namespace

size_t radix_sort_temp_bytes(unsigned int items);   // frontend.cu
size_t scan_temp_bytes(unsigned int items);         // predict_gp.cu
size_t lv_ray_info_bytes();                         // frontend_lv.cu
size_t lv_qgrid_bytes();                            // predict_lv.cu

Map::~Map() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (d_params) cudaFree(d_params);
    if (d_lut) cudaFree(d_lut);
    if (d_mm) cudaFree(d_mm);
    if (d_grid) cudaFree(d_grid);
    if (d_cnt) cudaFree(d_cnt);
    if (h_cnt) cudaFreeHost(h_cnt);
    if (d_args) cudaFree(d_args);
    if (d_peers) cudaFree(d_peers);
    if (fz_bar) cudaFree(fz_bar);
    if (h_args) cudaFreeHost(h_args);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_p0) cudaEventDestroy(ev_p0);
    if (ev_p1) cudaEventDestroy(ev_p1);
    if (ev_wait) cudaEventDestroy(ev_wait);
    if (stream) cudaStreamDestroy(stream);
}

void Map::init(int method, const la3dm_params &p, int dev) {
    if (method < LA3DM_BGK || method > LA3DM_GP) throw StatusError{LA3DM_ERR_INVALID, "unknown method"};
    if (p.block_depth < 1 || p.block_depth > kMaxDepth) throw StatusError{LA3DM_ERR_INVALID, "block_depth out of range"};
    if (!(p.resolution > 0) || !(p.ell > 0)) throw StatusError{LA3DM_ERR_INVALID, "resolution and ell must be > 0"};
    // configurations the predict kernels do not cover are refused HERE, before any scan can touch the map
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw StatusError{LA3DM_ERR_NO_DEVICE, "no CUDA device (la3dm_b200 has no CPU fallback)"};
    }
    if (dev < 0 || dev >= ndev) throw StatusError{LA3DM_ERR_INVALID, "device ordinal out of range"};
    device = dev;
    LA3DM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LA3DM_CUDA(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    LA3DM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    LA3DM_CUDA(cudaEventCreate(&ev0));
    LA3DM_CUDA(cudaEventCreate(&ev1));
    LA3DM_CUDA(cudaEventCreate(&ev_p0));
    LA3DM_CUDA(cudaEventCreate(&ev_p1));
    LA3DM_CUDA(cudaEventCreateWithFlags(&ev_wait, cudaEventDisableTiming));
    const char *env = getenv("LA3DM_NO_GRAPH");
    use_graph = !(env && env[0] == '1');
    const char *legacy = getenv("LA3DM_LEGACY_FRONTEND");
    use_fused = !(legacy && legacy[0] == '1');
    if (getenv("LA3DM_MEGA_TOT")) mega_tot = (unsigned int) std::max(32, atoi(getenv("LA3DM_MEGA_TOT")));
    if (getenv("LA3DM_MEGA_CHUNK")) mega_chunk = (unsigned int) std::max(32, atoi(getenv("LA3DM_MEGA_CHUNK")));
    if (mega_chunk > mega_tot) mega_chunk = mega_tot;

    api_params = p;
    DevParams &h = hp;
    std::memset(&h, 0, sizeof(h));
    h.method = method;
    h.depth = p.block_depth;
    int off = 0, n = 1;
    for (int d = 0; d < h.depth; ++d, n *= 8) { h.layer_off[d] = off; off += n; }
    h.layer_off[h.depth] = off;
    h.nodes = off;
    h.finest = n / 8;
    h.st_off = h.nodes * 8;
    h.rec_bytes = (h.nodes * 9 + 15) / 16 * 16;
    h.resolution = p.resolution;
    // block_size((float) pow(2, block_depth - 1) * resolution)   (src/bgkoctomap/bgkoctomap.cpp:41)
    h.block_size = (float) std::pow(2, p.block_depth - 1) * p.resolution;
    h.half_size = h.block_size / 2.0f;
    h.sf2 = p.sf2; h.ell = p.ell;
    h.free_thresh = p.free_thresh; h.occupied_thresh = p.occupied_thresh; h.var_thresh = p.var_thresh;
    h.prior_A = p.prior_A; h.prior_B = p.prior_B;
    h.min_W = p.min_W; h.original_size = p.original_size;
    h.noise = p.noise; h.l = p.l;
    if (method == LA3DM_GP) {   // src/gpoctomap/gpoctomap.cpp:38-40
        h.min_ivar = 1.0f / p.max_var; h.max_ivar = 1.0f / p.min_var; h.min_known_ivar = 1.0f / p.max_known_var;
        h.def_a = 0.0f; h.def_b = h.min_ivar;                       // gpoctree_node.h:34
    } else {
        h.def_a = p.prior_A; h.def_b = p.prior_B;                   // bgkoctree_node.h:34
    }
    h.pruned_state = method == LA3DM_BGKLV ? LA3DM_LV_PRUNED : LA3DM_PRUNED;

    // node-centre look-up table: init_key_loc_map (src/bgkoctomap/bgkblock.cpp:7-32), breadth first, child i of a
    // node gets +-quarter-edge offsets from bits 4 (x), 2 (y), 1 (z) of i, with the reference's float/double mix
    h_lut.assign(h.nodes, make_float3(0.f, 0.f, 0.f));
    for (int d = 0; d + 1 < h.depth; ++d) {
        const float half_size = (float) (p.resolution * std::pow(2, h.depth - d - 1) * 0.5f);
        const int cnt = 1 << (3 * d);
        for (int idx = 0; idx < cnt; ++idx) {
            const float3 c = h_lut[h.layer_off[d] + idx];
            for (int i = 0; i < 8; ++i) {
                float3 ch;
                ch.x = (float) (c.x + half_size * (i & 4 ? 0.5 : -0.5));
                ch.y = (float) (c.y + half_size * (i & 2 ? 0.5 : -0.5));
                ch.z = (float) (c.z + half_size * (i & 1 ? 0.5 : -0.5));
                h_lut[h.layer_off[d + 1] + idx * 8 + i] = ch;
            }
        }
    }
    if (h.depth == 3) {
        const int bit[3] = {4, 2, 1};
        for (int a = 0; a < 3; ++a)
            for (int j = 0; j < 7; ++j) {
                const int node = j < 4 ? 9 + 8 * bit[a] * (j >> 1) + bit[a] * (j & 1) : (j < 6 ? 1 + bit[a] * (j - 4) : 0);
                const float3 o = h_lut[node];
                h.ax_off[a][j] = a == 0 ? o.x : (a == 1 ? o.y : o.z);
            }
    }
    LA3DM_CUDA(cudaMalloc(&d_params, sizeof(DevParams)));
    LA3DM_CUDA(cudaMemcpy(d_params, &h, sizeof(DevParams), cudaMemcpyHostToDevice));
    LA3DM_CUDA(cudaMalloc(&d_lut, sizeof(float3) * h.nodes));
    LA3DM_CUDA(cudaMemcpy(d_lut, h_lut.data(), sizeof(float3) * h.nodes, cudaMemcpyHostToDevice));
    LA3DM_CUDA(cudaMalloc(&d_mm, sizeof(unsigned int) * 18));
    LA3DM_CUDA(cudaMalloc(&d_grid, sizeof(GridDesc)));
    LA3DM_CUDA(cudaMalloc(&d_cnt, sizeof(ScanCounters)));
    LA3DM_CUDA(cudaMemset(d_cnt, 0, sizeof(ScanCounters)));
    LA3DM_CUDA(cudaMallocHost(&h_cnt, sizeof(ScanCounters)));
    std::memset(h_cnt, 0, sizeof(ScanCounters));
    LA3DM_CUDA(cudaMalloc(&d_args, sizeof(ScanArgs)));
    LA3DM_CUDA(cudaMallocHost(&h_args, sizeof(ScanArgs)));
    std::memset(h_args, 0, sizeof(ScanArgs));
    // first guess of the per-scan capacities; they follow the scans from here on (grown on overflow, see insert_device)
    caps.points = 4096;
    caps.raw = 16 * caps.points;
    caps.train = caps.points + caps.raw;
    caps.members = caps.train + caps.train / 8;
    caps.cells = 1u << 16;
    caps.tests = 8192;
    caps.vg_cells = 1u << 20;
    caps.gp_store = method == LA3DM_GP ? (1u << 22) : 0;
    caps.gp_n_max = method == LA3DM_GP ? 160 : 0;
    caps.lv_active = method == LA3DM_BGKLV ? (1u << 18) : 0;
    ensure_pool(4096 + caps.tests);
    ensure_workspace();
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

// beam_sample walks a beam with  d = fr; while (d < l) { ...; d += fr; }  in fp32 (src/bgkoctomap/bgkoctomap.cpp:451-455):
// sample e sits at fr added to itself e times, whatever the hit.  The table restates that accumulation once per
// free_resolution so that the kernels index it instead of re-running the dependent additions for every hit.
constexpr unsigned int kBeamTab = 65536;

void Map::ensure_beam_table(float fr) {
    if (beam_tab.p && fr == beam_tab_fr) return;
    std::vector<float> t(kBeamTab);
    volatile float d = fr;
    for (unsigned int i = 0; i < kBeamTab; ++i) { t[i] = d; d = d + fr; }
    if (beam_tab.reserve(kBeamTab * sizeof(float), stream)) invalidate_graph();
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    LA3DM_CUDA(cudaMemcpy(beam_tab.p, t.data(), kBeamTab * sizeof(float), cudaMemcpyHostToDevice));
    beam_tab_fr = fr;
}

// timing events around the predict kernels: an event-record node when the scan is being captured into a graph
void Map::record_event(cudaEvent_t ev) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    LA3DM_CUDA(cudaStreamIsCapturing(stream, &st));
    if (st == cudaStreamCaptureStatusActive) LA3DM_CUDA(cudaEventRecordWithFlags(ev, stream, cudaEventRecordExternal));
    else LA3DM_CUDA(cudaEventRecord(ev, stream));
}

void Map::invalidate_graph() {
    if (graph_exec) {
        cudaStreamSynchronize(stream);
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
    }
}

// buffers for the current `caps` (grow-only); any reallocation invalidates the captured graph
void Map::ensure_workspace() {
    caps.train = caps.points + caps.raw;
    bool moved = false;
    const size_t n_sort = std::max<size_t>(std::max(caps.points, caps.raw), caps.members);
    for (int i = 0; i < 2; ++i) {
        moved |= sort_keys[i].reserve(n_sort * 4, stream);
        moved |= sort_vals[i].reserve(n_sort * 4, stream);
    }
    moved |= run_start.reserve((std::max<size_t>(caps.points, caps.raw) + 2) * 4, stream);
    moved |= tiles.reserve((n_sort / kTile + (size_t) caps.points / 256 + (size_t) caps.cells / 32 / 256 +
                            (size_t) caps.tests / 256 + 16) * 8, stream);
    moved |= long_list.reserve((size_t) 2 * (kMaxLongRuns + kMaxMidRuns) * 4, stream);
    moved |= long_flags.reserve((std::max<size_t>(caps.points, caps.raw) / 256 + 2) * 13, stream);
    moved |= hit_cnt.reserve((size_t) caps.points * 4, stream);
    moved |= hits_ds.reserve((size_t) caps.points * sizeof(float4), stream);
    moved |= frees_raw.reserve((size_t) caps.raw * sizeof(float4), stream);
    moved |= xy.reserve((size_t) caps.train * sizeof(float4), stream);
    moved |= pts_sorted.reserve((size_t) caps.members * sizeof(float4), stream);
    moved |= db_id.reserve(((size_t) caps.members + 2) * 4, stream);
    moved |= db_start.reserve(((size_t) caps.members + 2) * 4, stream);
    moved |= cell_db.reserve((size_t) caps.cells * 4, stream);
    moved |= test_bits.reserve(((size_t) caps.cells / 32 + 2) * 4, stream);
    moved |= test_id.reserve((size_t) caps.tests * 4, stream);
    moved |= plan.reserve((size_t) caps.tests * sizeof(NeighbourPlan), stream);
    moved |= heavy_list.reserve((size_t) caps.tests * 4, stream);
    moved |= light_list.reserve((size_t) caps.tests * 4, stream);
    if (hp.method == LA3DM_BGK) {
        // mega blocks (> kMegaTot neighbourhood points) and their chunks: sum of tot over all test blocks = 7 * members
        const size_t n_mega_cap = (size_t) 7 * caps.members / mega_tot + 8;
        const size_t n_chunk_cap = (size_t) 7 * caps.members / mega_chunk + n_mega_cap;
        moved |= mega_list.reserve(n_mega_cap * sizeof(uint4), stream);
        moved |= chunk_mega.reserve(n_chunk_cap * 4, stream);
        moved |= mega_acc.reserve(n_chunk_cap * 64 * sizeof(float2), stream);
    }
    if (hp.method == LA3DM_BGKLV) {
        moved |= lv_range.reserve((size_t) caps.points * 8, stream);
        moved |= lv_info.reserve((size_t) caps.points * lv_ray_info_bytes(), stream);
        moved |= ray_first.reserve((size_t) caps.points * 4, stream);
        moved |= lv_qgrid.reserve(lv_qgrid_bytes(), stream);
        moved |= lv_active.reserve((size_t) caps.lv_active * 8, stream);
        moved |= lv_blk_slot.reserve((size_t) caps.tests * 4, stream);
        moved |= lv_blk_flags.reserve((size_t) caps.tests, stream);
    }
    if (hp.method == LA3DM_BGKL || hp.method == LA3DM_BGKLV) {
        moved |= ray_of.reserve((size_t) caps.train * 4, stream);
        moved |= rays.reserve((size_t) caps.points * 2 * sizeof(float4), stream);
        moved |= segs.reserve((size_t) caps.members * 2 * sizeof(float4), stream);
        moved |= seg_start.reserve(((size_t) caps.members + 2) * 4, stream);
    }
    if (hp.method == LA3DM_GP || hp.method == LA3DM_BGKL) {
        // per-leaf results of one chunk of test blocks: chunk x 7 neighbours x leaves x 8 B (predict_gp.cu, predict_bgkl.cu)
        const size_t groups = (size_t) (hp.finest + 31) / 32, chunk = std::max<size_t>(1, 65536 / groups);
        moved |= gp_mv.reserve(std::min<size_t>(caps.tests, chunk) * 7 * groups * 32 * 8, stream);
    }
    if (hp.method == LA3DM_GP) {
        moved |= cell_test.reserve((size_t) caps.cells * 4, stream);
        moved |= gp_sizes.reserve(((size_t) caps.members + 2) * 8, stream);
        moved |= gp_off.reserve(((size_t) caps.members + 2) * 8, stream);
        moved |= gp_store.reserve((size_t) caps.gp_store * 4, stream);
        moved |= plan_db.reserve((size_t) caps.tests * 8 * 4, stream);
        gp_ctas = num_sms * 8;
        moved |= gp_scratch.reserve((size_t) gp_ctas * 4 * 2 * caps.gp_n_max * 32 * 4, stream);
    }
    const size_t tmp = std::max(radix_sort_temp_bytes((unsigned int) n_sort),
                                hp.method == LA3DM_GP ? scan_temp_bytes(caps.members) : (size_t) 0);
    if (tmp > cub_tmp_bytes) { moved |= cub_tmp.reserve(tmp, stream); cub_tmp_bytes = tmp; }
    if (hp.method == LA3DM_BGK || hp.method == LA3DM_GP) moved |= ensure_fused_workspace();
    if (hp.method == LA3DM_BGKL) moved |= ensure_bgkl_workspace();
    if (moved) invalidate_graph();
}

void Map::insert_device(const float *d_xyz, size_t n, size_t stride_bytes, const float origin[3], float ds, float fr,
                        float max_range, int mode) {
    const bool frontend_only = mode == 1, training = mode == 2;
    if (stride_bytes < (training ? 16u : 12u) || stride_bytes % 4 != 0)
        throw StatusError{LA3DM_ERR_INVALID, "stride_bytes must be a multiple of 4, >= 12 (16 for labelled points)"};
    if (training && hp.method != LA3DM_BGK && hp.method != LA3DM_GP)
        throw StatusError{LA3DM_ERR_UNSUPPORTED, "insert_training_data exists for BGKOctoMap and GPOctoMap only"};
    if (n > 0x7FFFFFF0ull) throw StatusError{LA3DM_ERR_INVALID, "too many points"};
    if (n > 0 && !d_xyz) throw StatusError{LA3DM_ERR_INVALID, "null cloud"};
    if (!training && !(fr > 0)) throw StatusError{LA3DM_ERR_INVALID, "free_res must be > 0"};
    if (!training && ds == 0) throw StatusError{LA3DM_ERR_INVALID, "ds_resolution must not be 0"};
    if (training) { ds = -1.f; fr = 1.f; }
    // BGKLV clamps the downsampling resolution to the map resolution (src/bgklvoctomap/bgklvoctomap.cpp:102-104)
    if (hp.method == LA3DM_BGKLV && ds > hp.resolution) ds = hp.resolution;
    LA3DM_CUDA(cudaSetDevice(device));
    d2h_bytes = 0;
    std::memset(&stats, 0, sizeof(stats));
    stats.n_points = (int64_t) n;
    last_T = 0;

    int call_replays = 0, call_captures = 0;
    for (int attempt = 0;; ++attempt) {
        if (attempt > 12) throw StatusError{LA3DM_ERR_NOMEM, "scan workspace did not settle"};
        if (n > caps.points) caps.points = grow_to((unsigned int) n, 4096);
        ensure_workspace();
        if (peers_attached && (size_t) n_blocks + caps.tests > pool_cap)
            throw StatusError{LA3DM_ERR_NOMEM, "the block pool would have to move while peers are attached: "
                                               "la3dm_reserve_blocks() more before la3dm_peer_attach()"};
        ensure_pool((size_t) n_blocks + caps.tests);
        ensure_beam_table(fr);

        ScanArgs &a = *h_args;
        a.xyz = d_xyz; a.n = (unsigned int) n; a.stride_f = (int) (stride_bytes / 4);
        a.ox = origin[0]; a.oy = origin[1]; a.oz = origin[2];
        a.ds = ds; a.inv_ds = 1.0f / ds; a.fr = fr; a.max_range = max_range;
        a.free_label = hp.method == LA3DM_GP ? -1.0f : 0.0f;   // src/gpoctomap/gpoctomap.cpp:399
        if (mode == 3) {
            // stage 1 of the ingest reads the transformed cloud with the prefilter's leaf size; k_ingest_commit then
            // switches the arguments to the prefiltered cloud and the scan's own ds_resolution
            if (stage_cloud.reserve((size_t) caps.points * sizeof(float4), stream)) invalidate_graph();
            a.raw_xyz = d_xyz; a.raw_stride_f = (int) (stride_bytes / 4);
            for (int q = 0; q < 12; ++q) a.tf[q] = ingest_tf[q];
            a.scan_ds = ds; a.scan_inv_ds = 1.0f / ds; a.min_points = ingest_min_points;
            a.stage_cloud = stage_cloud.as<float4>();
            a.xyz = stage_cloud.as<float>(); a.stride_f = 4;
            if (ingest_pre_ds > 0) { a.ds = ingest_pre_ds; a.inv_ds = 1.0f / ingest_pre_ds; }
        }
        a.frontend_only = frontend_only ? 1 : 0;
        a.training_data = training ? 1 : 0;
        a.shard_rank = shard_rank; a.shard_world = shard_world;
        a.n_blocks = (unsigned int) n_blocks; a.pool_cap = (unsigned int) pool_cap;
        a.beam_tab = beam_tab.as<float>(); a.beam_tab_n = kBeamTab;
        static const unsigned int heavy_tot = getenv("LA3DM_HEAVY_TOT") ? (unsigned int) atoi(getenv("LA3DM_HEAVY_TOT")) : kHeavyTot;
        a.heavy_tot = heavy_tot;
        a.mega_tot = mega_tot; a.mega_chunk = mega_chunk;
        a.peers = (peers_attached && !frontend_only) ? d_peers : nullptr;
        a.scan_seq = scan_seq + 1;
        static const int ab = getenv("LA3DM_AB") ? atoi(getenv("LA3DM_AB")) : 0;
        a.ab_flags = ab;

        LA3DM_CUDA(cudaEventRecord(ev0, stream));
        LA3DM_CUDA(cudaMemcpyAsync(d_args, h_args, sizeof(ScanArgs), cudaMemcpyHostToDevice, stream));
        if (use_graph) {
            if (!graph_exec || !(graph_caps == caps) || graph_mode != mode || graph_fused != fused_applicable(mode)) {
                invalidate_graph();
                cudaGraph_t g = nullptr;
                LA3DM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                try {
                    enqueue_scan(mode);
                } catch (...) {
                    cudaStreamEndCapture(stream, &g);
                    if (g) cudaGraphDestroy(g);
                    throw;
                }
                LA3DM_CUDA(cudaStreamEndCapture(stream, &g));
                const cudaError_t e = cudaGraphInstantiate(&graph_exec, g, 0);
                cudaGraphDestroy(g);
                LA3DM_CUDA(e);
                graph_caps = caps;
                graph_mode = mode;
                graph_fused = fused_applicable(mode);
                graph_launches = launches;
                ++call_captures;
            }
            LA3DM_CUDA(cudaGraphLaunch(graph_exec, stream));
            launches = graph_launches;
        } else {
            enqueue_scan(mode);
        }
        LA3DM_CUDA(cudaEventRecord(ev1, stream));
        // the one synchronisation of the scan: counters (and overflow bits) back to the host
        d2h_bytes += sizeof(ScanCounters);
        LA3DM_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(ScanCounters), cudaMemcpyDeviceToHost, stream));
        LA3DM_CUDA(cudaStreamSynchronize(stream));
        const unsigned int ovf = h_cnt->overflow;
        if (!ovf) break;
        // a workspace was too small: nothing was written to the map; grow from the sizes the device reports and replay
        ++replays;
        ++call_replays;
        if (ovf & OVF_EXTENT) throw StatusError{LA3DM_ERR_EXTENT, "scan bounding box spans too many blocks"};
        if (ovf & OVF_PEER) throw StatusError{LA3DM_ERR_CUDA, "timed out waiting for a peer replica to finish the scan"};
        if (ovf & OVF_FAST) use_fused = false;   // a list too long for the sort-free front-end: legacy pipeline from here on
        if (ovf & OVF_VGCELLS) {   // only the bit count matters (radix-sort passes): next power of two
            unsigned int v = 1u << 20;
            while (v < h_cnt->vg_cells_needed && v < 0x80000000u) v <<= 1;
            caps.vg_cells = v;
        }
        if (ovf & OVF_RAW) caps.raw = grow_to(h_cnt->n_raw_frees, 1024);
        if (ovf & OVF_CELLS) caps.cells = grow_to(h_cnt->n_cells, 1u << 16);
        if (ovf & OVF_MEMBERS) caps.members = grow_to(h_cnt->n_members, 1024);
        if (ovf & OVF_TESTS) caps.tests = grow_to(h_cnt->n_test_blocks, 8192);
        if (ovf & OVF_GPSTORE)
            caps.gp_store = grow_to((unsigned int) std::min<unsigned long long>(h_cnt->gp_store_needed, 0x60000000ull), 1u << 22);
        if (ovf & OVF_GPN) caps.gp_n_max = grow_to(h_cnt->gp_n_max, 160);
        if (ovf & OVF_LVACTIVE) caps.lv_active = grow_to(h_cnt->lv_active, 1u << 18);
        caps.train = caps.points + caps.raw;
        if (caps.members < caps.train / 2) caps.members = caps.train / 2;
    }

    dump_fused_trace();
    if (!frontend_only) {
        ++scan_seq;
        if (peers_attached && peers_deferred) peers_unsynced = true;
        static const bool wait_in_scan = getenv("LA3DM_PEER_WAIT_IN_SCAN") != nullptr;
        if (peers_attached && !peers_deferred && !wait_in_scan) peer_wait_pending = true;
    }
    // blocks after the scan = blocks before + blocks k_plan / k_lv_blocks created (no overflow on this path)
    n_blocks = frontend_only ? n_blocks : (long long) h_args->n_blocks + (long long) h_cnt->n_new_blocks;
    last_T = frontend_only ? 0 : h_cnt->n_test_blocks;
    stats.n_hits = h_cnt->n_hits;
    stats.n_train = h_cnt->n_train;
    stats.n_data_blocks = h_cnt->n_data_blocks;
    stats.n_test_blocks = h_cnt->n_test_blocks;
    stats.voxel_visits = (int64_t) h_cnt->visits;
    stats.voxel_updates = (int64_t) h_cnt->updates;
    stats.kernel_pairs = (int64_t) h_cnt->pairs;
    stats.n_blocks_total = n_blocks;
    stats.new_blocks = h_cnt->n_new_blocks;
    stats.kernel_launches = launches;
    stats.replays = call_replays;
    stats.graph_captures = call_captures;
    stats.grid_irregular = (int32_t) h_cnt->grid_irregular;
    stats.h2d_bytes = h2d_bytes + (long long) sizeof(ScanArgs);   // h2d_bytes: the cloud, set by the host entry point
    stats.d2h_bytes = d2h_bytes;
    h2d_bytes = 0;
    float ms = 0.f;
    LA3DM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    stats.device_ms = ms;
    if (!frontend_only && h_cnt->n_train > 0) {
        if (cudaEventElapsedTime(&ms, ev_p0, ev_p1) == cudaSuccess) stats.predict_ms = ms;
        else cudaGetLastError();
    }
}

// slots sorted by block key (deterministic export order)
void Map::sorted_block_order(DevBuf &order, size_t n) {
    for (int i = 0; i < 2; ++i) { order_keys[i].reserve(n * 8, stream); order_vals[i].reserve(n * 4, stream); }
    LA3DM_CUDA(cudaMemcpyAsync(order_keys[0].p, keys.p, n * 8, cudaMemcpyDeviceToDevice, stream));
    k_iota<<<ceil_div((long long) n, kThreads), kThreads, 0, stream>>>(order_vals[0].as<unsigned int>(), (unsigned int) n);
    cub::DoubleBuffer<long long> dk(order_keys[0].as<long long>(), order_keys[1].as<long long>());
    cub::DoubleBuffer<unsigned int> dv(order_vals[0].as<unsigned int>(), order_vals[1].as<unsigned int>());
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int) n, 0, 60, stream);
    export_tmp.reserve(tmp, stream);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(export_tmp.p, tmp, dk, dv, (int) n, 0, 60, stream));
    order.reserve(n * 4, stream);
    LA3DM_CUDA(cudaMemcpyAsync(order.p, dv.Current(), n * 4, cudaMemcpyDeviceToDevice, stream));
}

void Map::export_blocks(int64_t *out_keys, la3dm_node *out_nodes, size_t cap, size_t *n_out) {
    check_synced();
    LA3DM_CUDA(cudaSetDevice(device));
    const size_t n = (size_t) n_blocks;
    if (n_out) *n_out = n;
    if (n == 0 || (!out_keys && !out_nodes)) return;
    if (cap < n) throw StatusError{LA3DM_ERR_INVALID, "export_blocks: capacity too small"};
    sorted_block_order(block_order, n);
    const unsigned int *order = block_order.as<unsigned int>();
    if (out_keys) {
        export_buf.reserve(n * 8, stream);
        k_gather_keys<<<ceil_div((long long) n, kThreads), kThreads, 0, stream>>>(keys.as<long long>(), order,
                                                                                  (unsigned int) n,
                                                                                  export_buf.as<long long>());
        LA3DM_CUDA(cudaMemcpyAsync(out_keys, export_buf.p, n * 8, cudaMemcpyDeviceToHost, stream));
        LA3DM_CUDA(cudaStreamSynchronize(stream));
    }
    if (out_nodes) {
        const size_t total = n * (size_t) hp.nodes;
        export_buf.reserve(total * sizeof(la3dm_node), stream);
        k_pack_nodes<<<ceil_div((long long) total, kThreads), kThreads, 0, stream>>>(
            pool.as<unsigned char>(), order, (unsigned int) n, hp.nodes, hp.st_off, hp.rec_bytes,
            export_buf.as<la3dm_node>());
        LA3DM_CUDA(cudaMemcpyAsync(out_nodes, export_buf.p, total * sizeof(la3dm_node), cudaMemcpyDeviceToHost,
                                   stream));
        LA3DM_CUDA(cudaStreamSynchronize(stream));
    }
}

long long Map::count_leaves() {
    check_synced();
    LA3DM_CUDA(cudaSetDevice(device));
    const size_t n = (size_t) n_blocks;
    if (n == 0) return 0;
    sorted_block_order(block_order, n);
    const unsigned int *order = block_order.as<unsigned int>();
    leaf_cnt.reserve(n * 4, stream);
    leaf_off.reserve(n * 4, stream);
    k_leaf_count<<<ceil_div((long long) n * 32, kThreads), kThreads, 0, stream>>>(
        pool.as<unsigned char>(), order, (unsigned int) n, d_params, 0xFFu, leaf_cnt.as<unsigned int>());
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(), (int) n, stream);
    export_tmp.reserve(tmp, stream);
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(export_tmp.p, tmp, leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(),
                                             (int) n, stream));
    k_leaf_total<<<1, 1, 0, stream>>>(leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(), (unsigned int) n, d_cnt);
    unsigned int total = 0;
    LA3DM_CUDA(cudaMemcpyAsync(&total, &d_cnt->n_leaves, 4, cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    return total;
}

void Map::export_leaves(la3dm_leaf *out, size_t cap, size_t *n_out) {
    const long long total = count_leaves();   // leaves block_order (= key order) / leaf_off ready
    if (n_out) *n_out = (size_t) total;
    if (!out || total == 0) return;
    if (cap < (size_t) total) throw StatusError{LA3DM_ERR_INVALID, "export_leaves: capacity too small"};
    const size_t n = (size_t) n_blocks;
    leaf_out.reserve((size_t) total * sizeof(la3dm_leaf), stream);
    k_leaf_fill<<<ceil_div((long long) n * 32, kThreads), kThreads, 0, stream>>>(
        pool.as<unsigned char>(), keys.as<long long>(), block_order.as<unsigned int>(), (unsigned int) n, d_params,
        d_lut, leaf_off.as<unsigned int>(), 0xFFu, leaf_out.as<la3dm_leaf>());
    LA3DM_CUDA(cudaMemcpyAsync(out, leaf_out.p, (size_t) total * sizeof(la3dm_leaf), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace la3dm_b200

namespace la3dm_b200 {

// leaves (filtered by state) of the blocks order[0 .. n), in that order; out == nullptr: count only
void Map::leaves_of(const unsigned int *order, size_t n, unsigned int state_mask, la3dm_leaf *out, size_t cap, size_t *n_out) {
    leaf_cnt.reserve(n * 4, stream);
    leaf_off.reserve(n * 4, stream);
    k_leaf_count<<<ceil_div((long long) n * 32, kThreads), kThreads, 0, stream>>>(
        pool.as<unsigned char>(), order, (unsigned int) n, d_params, state_mask, leaf_cnt.as<unsigned int>());
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(), (int) n, stream);
    export_tmp.reserve(tmp, stream);
    LA3DM_CUDA(cub::DeviceScan::ExclusiveSum(export_tmp.p, tmp, leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(),
                                             (int) n, stream));
    k_leaf_total<<<1, 1, 0, stream>>>(leaf_cnt.as<unsigned int>(), leaf_off.as<unsigned int>(), (unsigned int) n, d_cnt);
    unsigned int total = 0;
    LA3DM_CUDA(cudaMemcpyAsync(&total, &d_cnt->n_leaves, 4, cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    if (n_out) *n_out = total;
    if (!out || total == 0) return;
    if (cap < (size_t) total) throw StatusError{LA3DM_ERR_INVALID, "leaf export: capacity too small"};
    leaf_out.reserve((size_t) total * sizeof(la3dm_leaf), stream);
    k_leaf_fill<<<ceil_div((long long) n * 32, kThreads), kThreads, 0, stream>>>(
        pool.as<unsigned char>(), keys.as<long long>(), order, (unsigned int) n, d_params, d_lut,
        leaf_off.as<unsigned int>(), state_mask, leaf_out.as<la3dm_leaf>());
    LA3DM_CUDA(cudaMemcpyAsync(out, leaf_out.p, (size_t) total * sizeof(la3dm_leaf), cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

// The server loop's mirror (bgkoctomap_server.cpp:94-144 walks the WHOLE map after every scan to rebuild its marker
// arrays): only the blocks that were test blocks of a scan since the last clearing call, only the leaves whose state is in
// state_mask.  block_keys lists every such block (sorted) -- a mirror replaces all it holds for those blocks, so a block
// whose wanted leaves have all gone (pruned, reclassified) is cleared too.
void Map::export_touched(unsigned int state_mask, la3dm_leaf *out, size_t cap, size_t *n_out, int64_t *block_keys,
                         size_t cap_blocks, size_t *n_blocks_out, bool clear) {
    check_synced();
    LA3DM_CUDA(cudaSetDevice(device));
    const size_t n = (size_t) n_blocks;
    if (n_out) *n_out = 0;
    if (n_blocks_out) *n_blocks_out = 0;
    if (n == 0) return;
    for (int i = 0; i < 2; ++i) { order_keys[i].reserve(n * 8, stream); order_vals[i].reserve(n * 4, stream); }
    unsigned int *d_count = &d_cnt->n_leaves;                                  // (scratch counter between scans)
    LA3DM_CUDA(cudaMemsetAsync(d_count, 0, 4, stream));
    k_touched_collect<<<ceil_div((long long) n, kThreads), kThreads, 0, stream>>>(
        touched.as<unsigned char>(), keys.as<long long>(), (unsigned int) n, order_keys[0].as<long long>(),
        order_vals[0].as<unsigned int>(), d_count);
    unsigned int nt = 0;
    LA3DM_CUDA(cudaMemcpyAsync(&nt, d_count, 4, cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    if (n_blocks_out) *n_blocks_out = nt;
    if (nt == 0) return;
    // by key: the order of the full export
    cub::DoubleBuffer<long long> dk(order_keys[0].as<long long>(), order_keys[1].as<long long>());
    cub::DoubleBuffer<unsigned int> dv(order_vals[0].as<unsigned int>(), order_vals[1].as<unsigned int>());
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int) nt, 0, 60, stream);
    export_tmp.reserve(tmp, stream);
    LA3DM_CUDA(cub::DeviceRadixSort::SortPairs(export_tmp.p, tmp, dk, dv, (int) nt, 0, 60, stream));
    block_order.reserve((size_t) nt * 4, stream);
    LA3DM_CUDA(cudaMemcpyAsync(block_order.p, dv.Current(), (size_t) nt * 4, cudaMemcpyDeviceToDevice, stream));
    if (block_keys) {
        if (cap_blocks < nt) throw StatusError{LA3DM_ERR_INVALID, "export_touched: block capacity too small"};
        LA3DM_CUDA(cudaMemcpyAsync(block_keys, dk.Current(), (size_t) nt * 8, cudaMemcpyDeviceToHost, stream));
    }
    leaves_of(block_order.as<unsigned int>(), nt, state_mask, out, cap, n_out);
    if (clear && out) LA3DM_CUDA(cudaMemsetAsync(touched.p, 0, n, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace la3dm_b200
