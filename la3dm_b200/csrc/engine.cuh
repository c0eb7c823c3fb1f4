// la3dm_b200 -- the device-resident map and the per-scan pipeline (host orchestration declared here).
#pragma once
#include "common.cuh"

namespace la3dm_b200 {

// One test block's view of its 7-block neighbourhood [self,+x,-x,+y,-y,+z,-z] (ExtendedBlock order,
// src/bgkoctomap/bgkblock.cpp:85-101): ranges into the block-sorted training array.
struct NeighbourPlan {
    unsigned int start[7];
    unsigned int count[7];
    unsigned int slot;      // block slot in the pool
    unsigned int is_new;    // created this scan: nodes still hold garbage, kernel writes defaults
};

struct Map {
    // ---- configuration
    la3dm_params api_params{};
    DevParams hp{};                 // host copy
    DevParams *d_params = nullptr;  // device copy
    float3 *d_lut = nullptr;        // [nodes] centre offset of each node (init_key_loc_map)
    std::vector<float3> h_lut;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_p0 = nullptr, ev_p1 = nullptr;
    int num_sms = 148;
    std::string last_error;

    // ---- persistent block pool (HBM-resident map)
    //   d_keys[slot]                   BlockHashKey
    //   d_ab  [slot * nodes + node]    float2 (m_A, m_B) | GP (m_ivar, ivar)
    //   d_st  [slot * nodes_pad + node] uint8: bits 0..2 state, bit 7 classified
    size_t pool_cap = 0;            // blocks
    int nodes_pad = 0;
    DevBuf keys, ab, st;
    long long n_blocks = 0;         // host mirror of *d_nblocks
    unsigned int *d_nblocks = nullptr;
    // open-addressing hash: key -> slot
    size_t hash_cap = 0;            // power of two
    DevBuf hkeys, hvals;

    // ---- per-scan workspace
    DevBuf cloud;                   // uploaded scan (host entry point)
    DevBuf sort_keys[2], sort_vals[2], flags, ranks, run_start, cub_tmp, scan64;
    DevBuf hits_ds;                 // float4 voxel-grid output of the cloud
    DevBuf frees_raw;               // float4 beam samples
    DevBuf xy;                      // float4 training set (x,y,z,label)
    DevBuf mem_cnt, mem_off;        // memberships per entry
    DevBuf pts_sorted;              // float4 block-sorted, pre-divided by ell
    DevBuf db_id, db_start;         // data blocks: dense id, start (+ sentinel)
    DevBuf cand[2], test_id, plan, miss;
    DevBuf shard_ids;
    unsigned int *d_mm = nullptr;   // [2][6] flipped min/max
    GridDesc *d_grid = nullptr;
    ScanCounters *d_cnt = nullptr;
    ScanCounters *h_cnt = nullptr;  // pinned
    la3dm_scan_stats stats{};
    int launches = 0;
    long long d2h_bytes = 0, h2d_bytes = 0;

    // ---- sharding (multi-GPU)
    int shard_rank = 0, shard_world = 1;
    unsigned int last_T = 0;        // test blocks of the last scan

    // ---- leaf export scratch
    DevBuf leaf_cnt, leaf_off, leaf_out, export_buf, order_keys[2], order_vals[2], block_order;

    Map() = default;
    ~Map();

    void init(int method, const la3dm_params &p, int device);
    void ensure_pool(size_t blocks);
    void insert_device(const float *d_xyz, size_t n, size_t stride_bytes, const float origin[3], float ds, float fr,
                       float max_range, bool frontend_only);
    // phases
    void frontend_bgk(const float *d_xyz, unsigned int n, int stride_f, float3 origin, float ds, float fr,
                      float max_range);
    unsigned int voxel_grid(const float *d_in, int stride_f, unsigned int n, float leaf, float4 *d_out,
                            const unsigned int *d_out_off, float label, unsigned int *d_count, int which);
    void minmax_points(const float *d_in, int stride_f, unsigned int n_host, const unsigned int *d_n,
                       unsigned int *mm, unsigned int n_upper);
    void bin_and_plan();
    void predict();
    void read_counters();
    // export
    void export_blocks(int64_t *keys, la3dm_node *nodes, size_t cap, size_t *n);
    long long count_leaves();
    void export_leaves(la3dm_leaf *out, size_t cap, size_t *n);
    void sorted_block_order(DevBuf &order, size_t n);
};

}  // namespace la3dm_b200

// the opaque C handle
struct la3dm_map {
    la3dm_b200::Map m;
};
