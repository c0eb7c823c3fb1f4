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
    unsigned int is_new;    // created this scan: the record still holds garbage, the predict kernel writes defaults
};
static_assert(sizeof(NeighbourPlan) == 64, "NeighbourPlan is loaded as four 16-byte words");

struct Map {
    // ---- configuration
    la3dm_params api_params{};
    DevParams hp{};                 // host copy
    DevParams *d_params = nullptr;  // device copy
    float3 *d_lut = nullptr;        // [nodes] centre offset of each node (init_key_loc_map)
    std::vector<float3> h_lut;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_p0 = nullptr, ev_p1 = nullptr, ev_wait = nullptr;
    int num_sms = 148;
    std::string last_error;

    // ---- persistent block pool (HBM-resident map).  One record per block, hp.rec_bytes bytes (a multiple of 16):
    //   float2 (m_A, m_B) | GP (m_ivar, ivar)  x nodes          at byte 0
    //   uint8  state (bits 0..2) | classified (bit 7)  x nodes  at byte hp.st_off
    // so that a block is one contiguous, 16-byte aligned span that a warp moves with 16-byte accesses.
    size_t pool_cap = 0;            // blocks
    DevBuf keys, pool;
    long long n_blocks = 0;
    // open-addressing hash: key -> slot
    size_t hash_cap = 0;            // power of two
    DevBuf hkeys, hvals;

    // ---- per-scan workspace, sized by `caps`
    Caps caps{};                    // logical capacities the scan kernels check against
    DevBuf cloud;                   // uploaded scan (host entry point)
    DevBuf stage_cloud;             // ingest: transformed / prefiltered cloud (float4)
    DevBuf sort_keys[2], sort_vals[2], run_start, cub_tmp, tiles, long_list, long_flags, hit_cnt;
    DevBuf hits_ds;                 // float4 voxel-grid output of the cloud
    DevBuf frees_raw;               // float4 beam samples
    DevBuf xy;                      // float4 training set (x,y,z,label)
    DevBuf pts_sorted;              // float4 block-sorted, pre-scaled for the method's kernel
    DevBuf db_id, db_start;         // data blocks: dense cell id, start (+ sentinel)
    DevBuf cell_db, test_bits;      // dense per-cell arrays of the scan's block grid
    DevBuf cell_test;               // GP: cell -> test block index + 1 (predict_gp_tc.cu walks data blocks)
    DevBuf test_id, plan, heavy_list, light_list, mega_list, chunk_mega, mega_acc;
    DevBuf gp_sizes, gp_off, gp_store, gp_scratch, gp_mv, plan_db;   // GPOctoMap: factor storage, per-leaf scratch
    int gp_ctas = 0;
    DevBuf bgkl_long_units, bgkl_chunk_unit, bgkl_partial, bgkl_long_cnt;   // BGKLOctoMap: long neighbour lists cut into chunks
    DevBuf ray_of, rays, segs, seg_start;   // BGKLOctoMap: ray of each marker, ray segments, per-block training lists
    DevBuf lv_range, lv_info, ray_first, lv_qgrid, lv_active, lv_blk_slot, lv_blk_flags;   // BGKLVOctoMap
    DevBuf fz_vcnt, fz_bits, fz_wpre, fz_cell_cnt, fz_extra, fz_tsum, fz_newsums, fz_bsum, fz_long;   // sort-free front-end (frontend_fused.cu)
    unsigned int *fz_bar = nullptr;
    unsigned int mega_tot = kMegaTot, mega_chunk = kMegaChunkPts;   // LA3DM_MEGA_TOT / LA3DM_MEGA_CHUNK
    bool use_fused = true;          // LA3DM_LEGACY_FRONTEND=1, or a scan that raised OVF_FAST, switches to frontend.cu / binning.cu
    DevBuf beam_tab;                // sample distances of beam_sample for the current free_resolution
    float beam_tab_fr = 0.f;
    size_t cub_tmp_bytes = 0;
    unsigned int *d_mm = nullptr;   // [3][6] flipped min/max
    GridDesc *d_grid = nullptr;
    ScanCounters *d_cnt = nullptr;
    ScanCounters *h_cnt = nullptr;  // pinned
    ScanArgs *d_args = nullptr;
    ScanArgs *h_args = nullptr;     // pinned
    la3dm_scan_stats stats{};
    int launches = 0;               // kernels enqueued by the last enqueue_scan()
    long long d2h_bytes = 0, h2d_bytes = 0;

    // ---- whole-scan CUDA graph (re-captured when capacities or buffers change)
    cudaGraphExec_t graph_exec = nullptr;
    Caps graph_caps{};
    int graph_mode = -1;
    bool graph_fused = false;
    int graph_launches = 0;
    bool use_graph = true;
    int replays = 0;                // scans re-run after a capacity overflow (lifetime counter)

    // ---- sharding (multi-GPU)
    int shard_rank = 0, shard_world = 1;
    unsigned int last_T = 0;        // test blocks of the last scan

    // ---- peers: replicas on other GPUs kept identical by direct stores from the predict kernel (shard.cu)
    PeerTable h_peers{};
    PeerTable *d_peers = nullptr;
    DevBuf peer_flags;              // [kMaxPeers] u64, written by the peers
    bool peer_wait_pending = false;    // eager mode: scans whose peer stores have not been waited for yet
    bool peers_share_device = false;   // a peer replica lives on this very device (single-GPU tests)
    bool peers_attached = false, peers_deferred = false, peers_unsynced = false;
    DevBuf touched;                 // [pool_cap] bytes: block was a test block of a scan since the last la3dm_export_touched(clear)
    DevBuf dirty;                   // [pool_cap] bytes: block changed since the last la3dm_peer_sync (deferred mode)
    unsigned long long sync_seq = 0;
    unsigned long long scan_seq = 0;

    // ---- leaf export scratch
    DevBuf leaf_cnt, leaf_off, leaf_out, export_buf, export_tmp, order_keys[2], order_vals[2], block_order;

    Map() = default;
    ~Map();

    void init(int method, const la3dm_params &p, int device);
    void ensure_pool(size_t blocks, bool exact = false);
    void ensure_workspace();
    void ensure_beam_table(float fr);
    void invalidate_graph();
    void record_event(cudaEvent_t ev);   // inside or outside of a stream capture
    // mode: 0 insert_pointcloud, 1 front-end only (get_training_data), 2 insert_training_data (d_xyz = x y z label),
    //       3 ingest (tf transform + prefilter from ingest_*) then insert_pointcloud
    void insert_device(const float *d_xyz, size_t n, size_t stride_bytes, const float origin[3], float ds, float fr,
                       float max_range, int mode);
    // the scan, enqueued on `stream` without host synchronisation (graph-capturable)
    void enqueue_scan(int mode);
    void enqueue_training_data();
    void enqueue_ingest();
    float ingest_tf[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    float ingest_pre_ds = -1.f;
    int ingest_min_points = 0;
    void enqueue_frontend_bgk();
    void enqueue_frontend_bgkl();
    void enqueue_bgkl_lists(const unsigned int *sorted_keys, const unsigned int *sorted_vals);
    void enqueue_predict_bgkl();
    void enqueue_frontend_lv();
    void enqueue_lv();
    void enqueue_block_grid();
    void enqueue_voxel_grid(int which);
    void enqueue_binning();
    void enqueue_predict();
    void enqueue_gp();
    void enqueue_gp_sizes();
    void enqueue_gp_mv_tc(unsigned int t0, unsigned int chunk);
    void enqueue_peer_wait();
    bool ensure_fused_workspace();
    bool ensure_bgkl_workspace();
    bool fused_applicable(int mode) const;
    void enqueue_fused_begin(int reset_counters);
    void enqueue_fused(int stage);
    void dump_fused_trace();
    void peer_sync();
    void check_synced() const;
    void peer_wait_now();
    // export
    void export_blocks(int64_t *keys, la3dm_node *nodes, size_t cap, size_t *n);
    long long count_leaves();
    void export_leaves(la3dm_leaf *out, size_t cap, size_t *n);
    void export_touched(unsigned int state_mask, la3dm_leaf *out, size_t cap, size_t *n, int64_t *block_keys,
                        size_t cap_blocks, size_t *n_blocks_out, bool clear);
    void leaves_of(const unsigned int *order, size_t n, unsigned int state_mask, la3dm_leaf *out, size_t cap, size_t *n_out);
    void sorted_block_order(DevBuf &order, size_t n);
    // query / import / serialisation (query.cu)
    void search(const float *xyz, size_t n, size_t stride_bytes, bool device_ptr, int finest_only, la3dm_leaf *out);
    void raycast(const float *start_end, size_t n_rays, size_t max_steps, la3dm_leaf *out, int32_t *n_steps);
    void import_blocks(const int64_t *keys, const la3dm_node *nodes, size_t n);
    void rebuild_hash();
    void save(const char *path);
    void load(const char *path);

    unsigned char *record(size_t slot) const { return pool.as<unsigned char>() + slot * (size_t) hp.rec_bytes; }
};

}  // namespace la3dm_b200

// the opaque C handle
struct la3dm_map {
    la3dm_b200::Map m;
};
