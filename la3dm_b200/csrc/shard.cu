// la3dm_b200 -- multi-GPU exchange of updated block states.
//
// Test blocks of a scan are dealt round-robin over ranks (t % world == rank) inside the fused predict kernel; every
// rank runs the front-end and the binning redundantly, so all ranks agree on the test-block list and the neighbour
// plan.  After the scan each rank packs the records of ITS test blocks into fixed-size rows; the caller all-gathers
// the rows over NCCL/NVLink and every rank scatters the peers' rows into its replica.
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

}  // namespace
}  // namespace la3dm_b200

using la3dm_b200::Map;

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

int la3dm_shard_pack(la3dm_map *map, void *d_rows) { return shard_copy(map, d_rows, 1); }
int la3dm_shard_unpack(la3dm_map *map, const void *d_all_rows) { return shard_copy(map, const_cast<void *>(d_all_rows), 0); }

}  // extern "C"
