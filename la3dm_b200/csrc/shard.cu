// la3dm_b200 -- multi-GPU exchange of updated block states.
//
// Test blocks of a scan are dealt round-robin over ranks (t % world == rank) inside the fused predict kernel; every
// rank runs the front-end and the binning redundantly, so all ranks agree on the test-block list and the neighbour
// plan.  After the scan each rank packs the records of ITS test blocks into fixed-size rows; the caller all-gathers
// the rows over NCCL/NVLink and every rank scatters the peers' rows into its replica.
#include <cstring>

#include <mutex>
#include <set>

#include "engine.cuh"

namespace la3dm_b200 {
namespace {

// row = one block record (hp.rec_bytes, a multiple of 16): (alpha, beta) pairs followed by the state bytes
__global__ void k_shard_copy(const NeighbourPlan *__restrict__ plan, const unsigned int *__restrict__ d_t,
                             unsigned char *pool, int rec_bytes, unsigned char *rows, unsigned int rows_per_rank,
                             int world, int my_rank, int pack) {
    const unsigned int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned int total_rows = pack ? rows_per_rank : rows_per_rank * (unsigned int) world;
    if (w >= total_rows) return;
    const unsigned int q = pack ? (unsigned int) my_rank : w / rows_per_rank;   // owner rank of this row
    const unsigned int r = pack ? w : w % rows_per_rank;
    if (!pack && (int) q == my_rank) return;
    const unsigned int t = r * (unsigned int) world + q;
    if (t >= *d_t) return;
    uint4 *row = reinterpret_cast<uint4 *>(rows + (size_t) w * rec_bytes);
    uint4 *rec = reinterpret_cast<uint4 *>(pool + (size_t) plan[t].slot * rec_bytes);
    const int words = rec_bytes >> 4;
    if (pack) for (int i = lane; i < words; i += 32) row[i] = rec[i];
    else for (int i = lane; i < words; i += 32) rec[i] = row[i];
}

// Peer replicas: waits until every attached peer has flagged this scan as pushed (its predict kernel's last CTA writes
// flags[rank] in our memory after a system-wide fence).  Bounded: a peer that never arrives raises OVF_PEER instead of
// hanging the device.
__global__ void k_peer_wait(const ScanArgs *__restrict__ A, ScanCounters *c, const unsigned long long *flags) {
    const PeerTable *PT = A->peers;
    if (!PT || c->overflow) return;
    const int p = threadIdx.x;
    if (p >= PT->world || p == PT->rank) return;
    const volatile unsigned long long *f = flags + p;
    const unsigned long long want = A->scan_seq;
    const long long t0 = clock64();
    while (*f < want) {
        if (clock64() - t0 > 20000000000ll) { atomicOr(&c->overflow, OVF_PEER); break; }   // ~10 s
        __nanosleep(200);
    }
    __threadfence_system();
}

// the same wait outside of a scan (read side): flags[p] >= want for every peer p
__global__ void k_peer_wait_lazy(const PeerTable *__restrict__ PT, const unsigned long long *flags, unsigned long long want,
                                 unsigned int *timed_out) {
    const int p = threadIdx.x;
    if (p >= PT->world || p == PT->rank) return;
    const volatile unsigned long long *f = flags + p;
    const long long t0 = clock64();
    while (*f < want) {
        if (clock64() - t0 > 20000000000ll) { *timed_out = 1u; break; }   // ~10 s
        __nanosleep(200);
    }
    __threadfence_system();
}

// Deferred mode: pushes every block this rank has changed since the last sync (only its owner ever marks a block) to all
// peers as whole records -- a warp per 32 slots, coalesced 16-byte stores -- and clears the marks; the last CTA then
// flags the sync as complete in every peer's memory.
__global__ void __launch_bounds__(256)
k_peer_sync(const PeerTable *__restrict__ PT, unsigned char *__restrict__ dirty, const unsigned char *__restrict__ pool,
            unsigned int n_blocks, int rec_bytes, unsigned long long sync_seq, unsigned int *done_counter) {
    const int lane = threadIdx.x & 31;
    const unsigned int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_w = (gridDim.x * blockDim.x) >> 5;
    const int world = PT->world, rank = PT->rank, words = rec_bytes >> 4;
    for (unsigned int base = gw * 32u; base < n_blocks; base += n_w * 32u) {
        const unsigned int slot = base + (unsigned int) lane;
        const bool d = slot < n_blocks && dirty[slot] != 0;
        unsigned int todo = __ballot_sync(0xffffffffu, d);
        if (d) dirty[slot] = 0;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const size_t off = (size_t) (base + (unsigned int) src) * (size_t) rec_bytes;
            for (int w = lane; w < words; w += 32) {
                const uint4 v = reinterpret_cast<const uint4 *>(pool + off)[w];
                for (int p = 0; p < world; ++p)
                    if (p != rank) reinterpret_cast<uint4 *>(PT->pool[p] + off)[w] = v;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done_counter, 1u) == gridDim.x - 1) {
        __threadfence_system();
        for (int p = 0; p < world; ++p)
            if (p != rank) *reinterpret_cast<volatile unsigned long long *>(PT->flags[p] + kMaxPeers + rank) = sync_seq;
        *done_counter = 0;
    }
}

__global__ void k_peer_sync_wait(const PeerTable *__restrict__ PT, const unsigned long long *flags,
                                 unsigned long long sync_seq, unsigned int *timed_out) {
    const int p = threadIdx.x;
    if (p >= PT->world || p == PT->rank) return;
    const volatile unsigned long long *f = flags + kMaxPeers + p;
    const long long t0 = clock64();
    while (*f < sync_seq) {
        if (clock64() - t0 > 40000000000ll) { *timed_out = 1; break; }     // ~20 s
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace

void Map::check_synced() const {
    if (peers_attached && peers_deferred && peers_unsynced)
        throw StatusError{LA3DM_ERR_INVALID, "this replica holds only its own blocks' latest state: la3dm_peer_sync() "
                                             "(on every rank) before reading the map"};
    // eager mode: the peers' stores of the last scans may still be on their way (see enqueue_peer_wait)
    if (peer_wait_pending) const_cast<Map *>(this)->peer_wait_now();
}

// Blocks until every peer has flagged every scan up to scan_seq as pushed into this replica.
void Map::peer_wait_now() {
    if (!peers_attached || peers_deferred) { peer_wait_pending = false; return; }
    LA3DM_CUDA(cudaSetDevice(device));
    unsigned int *scratch = reinterpret_cast<unsigned int *>(peer_flags.as<unsigned long long>() + 2 * kMaxPeers);
    LA3DM_CUDA(cudaMemsetAsync(scratch + 2, 0, 4, stream));
    k_peer_wait_lazy<<<1, 32, 0, stream>>>(d_peers, peer_flags.as<unsigned long long>(), scan_seq, scratch + 2);
    unsigned int timed_out = 0;
    LA3DM_CUDA(cudaMemcpyAsync(&timed_out, scratch + 2, 4, cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    if (timed_out) throw StatusError{LA3DM_ERR_CUDA, "timed out waiting for a peer replica to finish its scans"};
    peer_wait_pending = false;
}

// collective: every rank calls it; on return all replicas are identical
void Map::peer_sync() {
    if (!peers_attached || !peers_deferred) return;
    LA3DM_CUDA(cudaSetDevice(device));
    ++sync_seq;
    unsigned int *scratch = reinterpret_cast<unsigned int *>(peer_flags.as<unsigned long long>() + 2 * kMaxPeers);
    LA3DM_CUDA(cudaMemsetAsync(scratch + 1, 0, 4, stream));
    if (n_blocks > 0 || true)
        k_peer_sync<<<num_sms * 4, 256, 0, stream>>>(d_peers, dirty.as<unsigned char>(), pool.as<unsigned char>(),
                                                     (unsigned int) n_blocks, hp.rec_bytes, sync_seq, scratch);
    k_peer_sync_wait<<<1, 32, 0, stream>>>(d_peers, peer_flags.as<unsigned long long>(), sync_seq, scratch + 1);
    unsigned int timed_out = 0;
    LA3DM_CUDA(cudaMemcpyAsync(&timed_out, scratch + 1, 4, cudaMemcpyDeviceToHost, stream));
    LA3DM_CUDA(cudaStreamSynchronize(stream));
    if (timed_out) throw StatusError{LA3DM_ERR_CUDA, "la3dm_peer_sync: timed out waiting for a peer"};
    peers_unsynced = false;
}

// Eager mode.  A rank only ever predicts -- reads and writes -- the blocks it owns, so the next scan does not depend on the
// peers' stores of this one; only a READ of the map does.  By default the wait is therefore left to the read side
// (check_synced: export, search, ray casting, save, detach) and the peers' stores overlap the next scan's front-end;
// LA3DM_PEER_WAIT_IN_SCAN=1 waits at the end of every scan instead.
void Map::enqueue_peer_wait() {
    if (!peers_attached || peers_deferred) return;
    static const bool in_scan = getenv("LA3DM_PEER_WAIT_IN_SCAN") != nullptr;
    if (!in_scan) return;
    k_peer_wait<<<1, 32, 0, stream>>>(d_args, d_cnt, peer_flags.as<unsigned long long>());
    ++launches;
}

}  // namespace la3dm_b200

using la3dm_b200::Map;

static std::mutex g_ipc_mutex;
static std::set<const void *> g_ipc_pools;      // pool bases obtained from la3dm_peer_ipc_open

extern "C" {

int64_t la3dm_shard_row_bytes(const la3dm_map *map) { return map ? (int64_t) map->m.hp.rec_bytes : -1; }

int64_t la3dm_shard_rows(const la3dm_map *map) {
    if (!map) return -1;
    const int64_t T = map->m.last_T, W = map->m.shard_world;
    return (T + W - 1) / W;
}

static int shard_copy(la3dm_map *map, void *rows, int pack) {
    if (!map || !rows) return LA3DM_ERR_INVALID;
    Map &m = map->m;
    const unsigned int rpr = (unsigned int) la3dm_shard_rows(map);
    if (rpr == 0) return LA3DM_OK;
    if (cudaSetDevice(m.device) != cudaSuccess) return LA3DM_ERR_CUDA;
    const unsigned int total = pack ? rpr : rpr * (unsigned int) m.shard_world;
    la3dm_b200::k_shard_copy<<<la3dm_b200::ceil_div((long long) total * 32, 256), 256, 0, m.stream>>>(
        m.plan.as<la3dm_b200::NeighbourPlan>(), &m.d_cnt->n_test_blocks, m.pool.as<unsigned char>(), m.hp.rec_bytes,
        static_cast<unsigned char *>(rows), rpr, m.shard_world, m.shard_rank, pack);
    // stream-ordered on la3dm_stream(map): the caller orders its collective after / before these kernels on that stream
    if (cudaGetLastError() != cudaSuccess) {
        m.last_error = "shard copy kernel launch failed";
        return LA3DM_ERR_CUDA;
    }
    return LA3DM_OK;
}

// ---- peer replicas ----------------------------------------------------------------------------------------------------
static int peer_fail(la3dm_map *map, const char *what, cudaError_t e) {
    map->m.last_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return LA3DM_ERR_CUDA;
}

static int ensure_peer_buffers(la3dm_map *map) {
    Map &m = map->m;
    if (cudaSetDevice(m.device) != cudaSuccess) return LA3DM_ERR_CUDA;
    if (!m.peer_flags.p) {
        try { m.peer_flags.reserve((2 * la3dm_b200::kMaxPeers + 2) * sizeof(unsigned long long), m.stream); }
        catch (...) { return LA3DM_ERR_NOMEM; }
        cudaMemsetAsync(m.peer_flags.p, 0, m.peer_flags.cap, m.stream);
        cudaStreamSynchronize(m.stream);
    }
    if (!m.d_peers && cudaMalloc(&m.d_peers, sizeof(la3dm_b200::PeerTable)) != cudaSuccess) return LA3DM_ERR_NOMEM;
    return LA3DM_OK;
}

int la3dm_reserve_blocks(la3dm_map *map, size_t blocks) {
    if (!map) return LA3DM_ERR_INVALID;
    if (map->m.peers_attached) { map->m.last_error = "reserve_blocks: detach the peers first"; return LA3DM_ERR_INVALID; }
    try {
        cudaSetDevice(map->m.device);
        map->m.ensure_pool(blocks, true);
        cudaStreamSynchronize(map->m.stream);
    } catch (const la3dm_b200::CudaError &e) { return peer_fail(map, "reserve_blocks", e.code); }
    catch (...) { return LA3DM_ERR_NOMEM; }
    return LA3DM_OK;
}

int la3dm_peer_local(la3dm_map *map, void **pool_base, void **flags) {
    if (!map || !pool_base || !flags) return LA3DM_ERR_INVALID;
    const int rc = ensure_peer_buffers(map);
    if (rc != LA3DM_OK) return rc;
    *pool_base = map->m.pool.p;
    *flags = map->m.peer_flags.p;
    return LA3DM_OK;
}

int la3dm_peer_ipc_export(la3dm_map *map, void *handle_pool, void *handle_flags) {
    if (!map || !handle_pool || !handle_flags) return LA3DM_ERR_INVALID;
    const int rc = ensure_peer_buffers(map);
    if (rc != LA3DM_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == LA3DM_IPC_HANDLE_BYTES, "handle size");
    cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle_pool), map->m.pool.p);
    if (e != cudaSuccess) return peer_fail(map, "cudaIpcGetMemHandle(pool)", e);
    e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle_flags), map->m.peer_flags.p);
    if (e != cudaSuccess) return peer_fail(map, "cudaIpcGetMemHandle(flags)", e);
    return LA3DM_OK;
}

int la3dm_peer_ipc_open(la3dm_map *map, const void *handle_pool, const void *handle_flags, void **pool_base, void **flags) {
    if (!map || !handle_pool || !handle_flags || !pool_base || !flags) return LA3DM_ERR_INVALID;
    if (cudaSetDevice(map->m.device) != cudaSuccess) return LA3DM_ERR_CUDA;
    cudaIpcMemHandle_t hp_, hf_;
    memcpy(&hp_, handle_pool, sizeof(hp_));
    memcpy(&hf_, handle_flags, sizeof(hf_));
    cudaError_t e = cudaIpcOpenMemHandle(pool_base, hp_, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return peer_fail(map, "cudaIpcOpenMemHandle(pool)", e);
    e = cudaIpcOpenMemHandle(flags, hf_, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return peer_fail(map, "cudaIpcOpenMemHandle(flags)", e);
    {   // a pool opened through IPC belongs to another process (cudaPointerGetAttributes reports the MAPPING device for it)
        std::lock_guard<std::mutex> lock(g_ipc_mutex);
        g_ipc_pools.insert(*pool_base);
    }
    return LA3DM_OK;
}

int la3dm_peer_attach(la3dm_map *map, int world, int rank, void *const *pool_bases, void *const *flags) {
    if (!map || world < 2 || world > la3dm_b200::kMaxPeers || rank < 0 || rank >= world || !pool_bases || !flags)
        return LA3DM_ERR_INVALID;
    Map &m = map->m;
    if (m.hp.method != LA3DM_BGK || m.hp.depth > 3) {
        m.last_error = "peer replicas are implemented for BGKOctoMap with block_depth <= 3";
        return LA3DM_ERR_UNSUPPORTED;
    }
    const int rc = ensure_peer_buffers(map);
    if (rc != LA3DM_OK) return rc;
    la3dm_b200::PeerTable &t = m.h_peers;
    memset(&t, 0, sizeof(t));
    t.world = world; t.rank = rank; t.deferred = m.peers_deferred ? 1 : 0;
    for (int p = 0; p < world; ++p) {
        if (p != rank && (!pool_bases[p] || !flags[p])) return LA3DM_ERR_INVALID;
        t.pool[p] = p == rank ? m.pool.as<unsigned char>() : static_cast<unsigned char *>(pool_bases[p]);
        t.flags[p] = p == rank ? m.peer_flags.as<unsigned long long>() : static_cast<unsigned long long *>(flags[p]);
    }
    cudaStreamSynchronize(m.stream);
    cudaError_t e = cudaMemcpy(m.d_peers, &t, sizeof(t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return peer_fail(map, "peer table upload", e);
    m.peers_share_device = false;
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        cudaPointerAttributes at;
        {
            std::lock_guard<std::mutex> lock(g_ipc_mutex);
            if (g_ipc_pools.count(pool_bases[p])) continue;                    // another process: its own device
        }
        if (cudaPointerGetAttributes(&at, pool_bases[p]) == cudaSuccess) { if (at.device == m.device) m.peers_share_device = true; }
        else cudaGetLastError();
    }
    m.shard_rank = rank;
    m.shard_world = world;
    m.peers_attached = true;
    m.invalidate_graph();
    return LA3DM_OK;
}

int la3dm_peer_set_deferred(la3dm_map *map, int deferred) {
    if (!map) return LA3DM_ERR_INVALID;
    if (map->m.peers_attached) { map->m.last_error = "peer_set_deferred: call it before la3dm_peer_attach"; return LA3DM_ERR_INVALID; }
    map->m.peers_deferred = deferred != 0;
    return LA3DM_OK;
}

int la3dm_peer_sync(la3dm_map *map) {
    if (!map) return LA3DM_ERR_INVALID;
    try { map->m.peer_sync(); }
    catch (const la3dm_b200::CudaError &e) { return peer_fail(map, "peer_sync", e.code); }
    catch (const la3dm_b200::StatusError &e) { map->m.last_error = e.msg; return e.status; }
    return LA3DM_OK;
}

int la3dm_peer_detach(la3dm_map *map) {
    if (!map) return LA3DM_ERR_INVALID;
    if (map->m.peers_unsynced) { map->m.last_error = "peer_detach: la3dm_peer_sync() first"; return LA3DM_ERR_INVALID; }
    try {
        if (map->m.peer_wait_pending) map->m.peer_wait_now();      // the peers' last stores must have landed
    } catch (...) {
        map->m.last_error = "peer_detach: timed out waiting for a peer replica";
        return LA3DM_ERR_CUDA;
    }
    map->m.peers_attached = false;
    map->m.shard_rank = 0;
    map->m.shard_world = 1;
    map->m.invalidate_graph();
    return LA3DM_OK;
}

int la3dm_shard_pack(la3dm_map *map, void *d_rows) { return shard_copy(map, d_rows, 1); }
int la3dm_shard_unpack(la3dm_map *map, const void *d_all_rows) { return shard_copy(map, const_cast<void *>(d_all_rows), 0); }

}  // extern "C"
