/*
 * la3dm_b200 -- C ABI of the B200-native per-scan Bayesian kernel occupancy update.
 *
 * This header is the drop-in boundary for ONE path of RobustFieldAutonomyLab/la3dm: the insert_pointcloud() call of
 * BGKOctoMap / BGKLOctoMap / BGKLVOctoMap / GPOctoMap and the read side the ROS nodes use right after it.  The
 * reference has no FFI layer of its own (it is a C++ class called from the nodes); each entry point below cites the
 * reference member it replaces.  The C++ facade in include/la3dm_b200/ (same class and method names as the reference)
 * forwards to these functions; INTEGRATION.md shows how a maintainer wires it into the existing nodes.
 *
 * Plain C: pointers + sizes only, no C++/torch types, no exceptions across the boundary.  All functions return
 * LA3DM_OK (0) or a negative la3dm_status; la3dm_last_error() gives the text for the failing call on that map.
 * A map is bound to one CUDA device and must be used from one host thread at a time (the reference's callers are
 * single-threaded: src/bgkoctomap/bgkoctomap_server.cpp:195-204).
 */
#ifndef LA3DM_B200_H
#define LA3DM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LA3DM_B200_ABI_VERSION 1

typedef struct la3dm_map la3dm_map; /* opaque handle */

typedef enum la3dm_method {
    LA3DM_BGK = 0,   /* la3dm::BGKOctoMap   (include/bgkoctomap/bgkoctomap.h)    */
    LA3DM_BGKL = 1,  /* la3dm::BGKLOctoMap  (include/bgkloctomap/bgkloctomap.h)  */
    LA3DM_BGKLV = 2, /* la3dm::BGKLVOctoMap (include/bgklvoctomap/bgklvoctomap.h) */
    LA3DM_GP = 3     /* la3dm::GPOctoMap    (include/gpoctomap/gpoctomap.h)      */
} la3dm_method;

typedef enum la3dm_status {
    LA3DM_OK = 0,
    LA3DM_ERR_INVALID = -1,     /* bad argument                                        */
    LA3DM_ERR_CUDA = -2,        /* CUDA runtime failure (see la3dm_last_error)         */
    LA3DM_ERR_UNSUPPORTED = -3, /* method / parameter combination not implemented      */
    LA3DM_ERR_EXTENT = -4,      /* scan bounding box too large for the block-id space  */
    LA3DM_ERR_NOMEM = -5,
    LA3DM_ERR_NO_DEVICE = -6    /* no CUDA device: there is NO CPU fallback by design   */
} la3dm_status;

/* Node states: enum class State in include/bgkoctomap/bgkoctree_node.h:10-12.  BGKLV inserts UNCERTAIN before PRUNED
 * (include/bgklvoctomap/bgklvoctree_node.h:11-13); la3dm_export_* always reports the numbering of the map's method. */
enum { LA3DM_FREE = 0, LA3DM_OCCUPIED = 1, LA3DM_UNKNOWN = 2, LA3DM_PRUNED = 3 };
enum { LA3DM_LV_UNCERTAIN = 3, LA3DM_LV_PRUNED = 4 };

/* Constructor arguments of the four reference map classes (the reference keeps them in class statics:
 * src/bgkoctomap/bgkoctomap.cpp:42-55).  Unused fields for a method are ignored. */
typedef struct la3dm_params {
    float resolution;        /* all: finest voxel edge (m)                                            */
    int32_t block_depth;     /* all: test-data octree depth; block edge = 2^(depth-1) * resolution   */
    float sf2;               /* all: kernel signal variance                                           */
    float ell;               /* all: kernel length scale                                              */
    float free_thresh;       /* all                                                                   */
    float occupied_thresh;   /* all                                                                   */
    float var_thresh;        /* BGK, BGKL, BGKLV                                                      */
    float prior_A;           /* BGK, BGKL, BGKLV                                                      */
    float prior_B;           /* BGK, BGKL, BGKLV                                                      */
    int32_t original_size;   /* BGKLV (bool)                                                          */
    float min_W;             /* BGKLV                                                                 */
    float noise;             /* GP                                                                    */
    float l;                 /* GP                                                                    */
    float min_var;           /* GP                                                                    */
    float max_var;           /* GP                                                                    */
    float max_known_var;     /* GP                                                                    */
} la3dm_params;

/* One octree node in the reference's own in-memory layout: class Occupancy { bool classified; float m_A; float m_B;
 * State state; } = 16 bytes (include/bgkoctomap/bgkoctree_node.h:76-81).  For GP the floats are (m_ivar, ivar)
 * (include/gpoctomap/gpoctree_node.h). */
typedef struct la3dm_node {
    uint8_t classified;
    uint8_t _pad0[3];
    float a;
    float b;
    uint8_t state;
    uint8_t _pad1[3];
} la3dm_node;

/* One leaf as LeafIterator exposes it (include/bgkoctomap/bgkoctomap.h:217-307: get_loc / get_size / get_node). */
typedef struct la3dm_leaf {
    int64_t block_key;  /* BlockHashKey (src/bgkoctomap/bgkblock.cpp:73-77)                      */
    int32_t depth;      /* node depth in the block's octree, 0 = whole block                     */
    int32_t index;      /* node index within its layer (child i of node k is 8k+i)               */
    float x, y, z;      /* centre (Block::get_loc)                                               */
    float size;         /* edge length (Block::get_size)                                         */
    float a, b;         /* (m_A, m_B) or, for GP, (m_ivar, ivar)                                 */
    float prob;         /* Occupancy::get_prob()                                                 */
    float var;          /* Occupancy::get_var()                                                  */
    uint8_t state;
    uint8_t classified;
    uint8_t _pad[6];    /* sizeof(la3dm_leaf) == 56 */
} la3dm_leaf;

/* Counters of the last insert call (the reference only prints some of these under -DDEBUG:
 * src/bgkoctomap/bgkoctomap.cpp:226,286-287). */
typedef struct la3dm_scan_stats {
    int64_t n_points;        /* input points                                                         */
    int64_t n_hits;          /* hits kept after downsampling and the range filter                    */
    int64_t n_train;         /* training entries (hits + free samples / markers)                     */
    int64_t n_data_blocks;   /* blocks that hold >= 1 training entry                                 */
    int64_t n_test_blocks;   /* blocks predicted this scan                                           */
    int64_t voxel_visits;    /* leaves of test blocks processed                                      */
    int64_t voxel_updates;   /* visits for which Occupancy::update fired at least once               */
    int64_t kernel_pairs;    /* (leaf, training entry) kernel evaluations the reference would do     */
    int64_t n_blocks_total;  /* blocks in the map after the call                                     */
    int64_t new_blocks;      /* blocks created by this call                                          */
    int32_t kernel_launches; /* GPU kernels launched by this call (ours + CUB primitives)            */
    int32_t grid_irregular;  /* 1 if the float-stepped block grid skipped or repeated an index       */
    float device_ms;         /* device time of the call, CUDA events on the map's stream             */
    float predict_ms;        /* device time of the fused predict/update/prune kernel alone           */
    int64_t h2d_bytes;       /* bytes copied host -> device by this call (the cloud)                  */
    int64_t d2h_bytes;       /* bytes copied device -> host by this call (counters)                   */
    int32_t replays;         /* times this call re-ran the scan after growing a workspace (map untouched) */
    int32_t graph_captures;  /* 1 if this call (re)captured the scan's CUDA graph                     */
} la3dm_scan_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
/* Replaces the map constructors (include/bgkoctomap/bgkoctomap.h:50-58, bgkloctomap.h:53-61, bgklvoctomap.h:52-62,
 * gpoctomap.h:50-52).  device = CUDA ordinal.  Fails with LA3DM_ERR_NO_DEVICE when no GPU is present. */
int la3dm_create(int method, const la3dm_params *params, int device, la3dm_map **out);
int la3dm_destroy(la3dm_map *map);
const char *la3dm_last_error(const la3dm_map *map);
const char *la3dm_status_string(int status);
int la3dm_abi_version(void);

/* ---- the hot path -------------------------------------------------------------------------------------------- */
/* Replaces  void insert_pointcloud(const PCLPointCloud &cloud, const point3f &origin, float ds_resolution,
 *                                  float free_res = 2.0f, float max_range = -1)
 * (include/bgkoctomap/bgkoctomap.h:82-84; src/bgkoctomap/bgkoctomap.cpp:214-366 and the -L/-LV/GP copies).
 * xyz: HOST pointer to n points, x y z as float32 at the start of each stride_bytes-sized record (12 for packed xyz,
 * 16 for pcl::PointXYZ).  The cloud is not retained.  Synchronous: the map is up to date on return.
 * The cloud may live in pageable memory (pcl's cloud.points does; the copy is then staged by the CUDA driver) or in
 * pinned memory (cudaHostAlloc / cudaHostRegister): 0.68 vs 0.62 ms per 64 k-point scan end to end (bench.py, "e2e").
 * With peer replicas attached (la3dm_peer_attach, eager mode) "up to date" holds for the blocks this replica owns; the
 * peers' blocks are waited for by the first read-side call (export, search, ray casting, save, detach). */
int la3dm_insert_pointcloud(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float origin[3],
                            float ds_resolution, float free_res, float max_range);

/* Same, with the scan already resident in device memory of the map's device (no host<->device copy of the cloud).
 * Work is enqueued on the map's OWN stream (la3dm_stream, created non-blocking); returns after the scan has been fully
 * applied.  The cloud must be complete before the call: if another stream is still writing it (a kernel, an async
 * copy), either synchronise that stream or call la3dm_stream_wait(map, that_stream) first. */
int la3dm_insert_pointcloud_device(la3dm_map *map, const float *d_xyz, size_t n, size_t stride_bytes,
                                   const float origin[3], float ds_resolution, float free_res, float max_range);

/* The live node's steps in front of insert_pointcloud, fused into the GPU front-end (cloudHandler,
 * src/bgkoctomap/bgkoctomap_server.cpp:70-86 and the -L/GP copies; the -LV server skips the prefilter,
 * bgklvoctomap_server.cpp:76-77): the sensor-frame cloud is moved into the map frame by tf (pcl_ros::
 * transformPointCloud; 3 x 4 row-major, fp32, evaluated like pcl::transformPointCloud), downsampled by a
 * pcl::VoxelGrid of leaf prefilter_ds (<= 0: no prefilter), and inserted only if more than min_points points are left
 * (upstream: 5) -- otherwise the call is a no-op, like upstream's `if (filtered_cloud.size() > 5)`.  origin is the
 * translation of the transform, as upstream passes it; ds_resolution / free_res / max_range are insert_pointcloud's.
 * The TF lookup and the motion gate (:46-60) stay with the caller: they are ROS calls, not arithmetic on the cloud. */
int la3dm_insert_pointcloud_ingest(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float tf[12],
                                   float prefilter_ds, int min_points, const float origin[3], float ds_resolution,
                                   float free_res, float max_range);

/* Replaces  void insert_training_data(const GPPointCloud &xy)  (include/bgkoctomap/bgkoctomap.h:86,
 * src/bgkoctomap/bgkoctomap.cpp:82-212; include/gpoctomap/gpoctomap.h, src/gpoctomap/gpoctomap.cpp:71-203): the same
 * update WITHOUT the front-end -- the caller's pre-labelled points are the training set.  xyzy: n records of
 * stride_bytes (>= 16), x y z label as float32.  As upstream, BGKOctoMap applies Occupancy::update to every leaf of every
 * test block here (bgkoctomap.cpp:179 has no `kbar > 0` guard: untouched voxels become `classified`).  Upstream
 * dereferences a null Block* when a test block does not exist yet (:155-160, `block` stays nullptr after emplace); here
 * the block is created, like insert_pointcloud does.  BGKOctoMap and GPOctoMap only (the -L/-LV classes have no such
 * member): LA3DM_ERR_UNSUPPORTED otherwise. */
int la3dm_insert_training_data(la3dm_map *map, const float *xyzy, size_t n, size_t stride_bytes);
int la3dm_insert_training_data_device(la3dm_map *map, const float *d_xyzy, size_t n, size_t stride_bytes);

/* Front-end only: get_training_data() (src/bgkoctomap/bgkoctomap.cpp:383-417).  Runs the GPU front-end on a HOST
 * cloud and returns the training set; does not touch the map.  Call with out == NULL to get the count.
 * out: 7 floats per entry (x0 y0 z0 x1 y1 z1 label); for BGK/GP x1..z1 repeat x0..z0. */
int la3dm_training_data(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float origin[3],
                        float ds_resolution, float free_res, float max_range, float *out, size_t capacity,
                        size_t *n_out);

/* BGKL / BGKLV only: the ray segments and the ray index of every training entry produced by the LAST
 * la3dm_training_data / la3dm_insert_pointcloud call (get_training_data's `rays` and `ray_idx` outputs,
 * src/bgkloctomap/bgkloctomap.cpp:285-344, src/bgklvoctomap/bgklvoctomap.cpp:303-423).  rays: 6 floats per ray
 * (x0 y0 z0 x1 y1 z1); ray_idx: one int32 per training entry (-1 for a hit).  Either output may be NULL. */
int la3dm_training_rays(la3dm_map *map, float *rays, size_t ray_capacity, size_t *n_rays, int32_t *ray_idx,
                        size_t idx_capacity);

int la3dm_last_stats(const la3dm_map *map, la3dm_scan_stats *out);

/* ---- read side ----------------------------------------------------------------------------------------------- */
int64_t la3dm_num_blocks(const la3dm_map *map);
int32_t la3dm_nodes_per_block(const la3dm_map *map); /* (8^depth - 1) / 7 */

/* Whole map in the reference's Block/OcTree layout, for the host mirror behind begin_leaf()/search():
 * keys[i] = BlockHashKey; nodes[i * nodes_per_block + off(d) + index], off(d) = (8^d - 1)/7, is node_arr[d][index]
 * (src/bgkoctomap/bgkoctree.cpp:18-27).  Blocks come sorted by key. */
int la3dm_export_blocks(la3dm_map *map, int64_t *keys, la3dm_node *nodes, size_t capacity_blocks, size_t *n_blocks);

/* All current leaves, compacted on the GPU (what the nodes walk with begin_leaf()..end_leaf():
 * src/bgkoctomap/bgkoctomap_static_node.cpp:111-139).  Sorted by (block_key, depth, index). */
int64_t la3dm_num_leaves(la3dm_map *map);
int la3dm_export_leaves(la3dm_map *map, la3dm_leaf *out, size_t capacity, size_t *n_out);

/* Incremental mirror for the server loop.  The reference's server re-walks the WHOLE map after every scan to rebuild
 * its marker arrays and skips everything that is not OCCUPIED / FREE (src/bgkoctomap/bgkoctomap_server.cpp:94-144).
 * This returns, compacted on the GPU, only the leaves whose state is in state_mask (bit s = state s, e.g.
 * (1 << LA3DM_OCCUPIED) | (1 << LA3DM_FREE); 0xFF = all) of the blocks that were test blocks of a scan (or were
 * imported / loaded) since the last call with clear != 0, sorted like la3dm_export_leaves, and the sorted keys of those
 * blocks: the mirror REPLACES what it holds for every listed block (a block whose wanted leaves have all gone is listed
 * with no leaves).  leaves == NULL: counts only (nothing is cleared); block_keys may be NULL. */
int la3dm_export_touched(la3dm_map *map, unsigned int state_mask, la3dm_leaf *leaves, size_t capacity_leaves,
                         size_t *n_leaves, int64_t *block_keys, size_t capacity_blocks, size_t *n_blocks, int clear);

/* Point query, batched: replaces  OcTreeNode search(point3f p) / search(float x, float y, float z)
 * (include/bgkoctomap/bgkoctomap.h:315-319; src/bgkoctomap/bgkoctomap.cpp:554-574 -> Block::search,
 * src/bgkoctomap/bgkblock.cpp:132-156).  xyz: HOST pointer, n points, stride_bytes per record; out: n HOST records.
 * out[i] describes the node that holds point i: block_key = block_to_hash_key(p); a block that does not exist answers
 * like upstream's `return OcTreeNode()` (default node) with depth = index = -1.  finest_only != 0 returns the
 * finest-layer node of the point's cell even when it is PRUNED (what upstream's operator[] hands back); 0 returns the
 * LEAF containing the point.  The cell index uses cell_num = 2^(block_depth-1): upstream freezes Block::cell_num at 8
 * during static initialisation (bgkblock.cpp:105), which is only right for block_depth 4. */
int la3dm_search(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, int finest_only, la3dm_leaf *out);

/* Ray casting, batched: replaces  class RayCaster { RayCaster(map, start, end); bool end(); bool next(p, node, block_key,
 * node_key); }  (include/bgkoctomap/bgkoctomap.h:91-214 and the -L/-LV/GP copies) -- the integer walk over the finest
 * cells from start to end, crossing blocks.  start_end: n_rays HOST records of 6 floats (start xyz, end xyz).  Step i of
 * ray r lands in out[r * max_steps + i] (HOST), n_steps[r] of them (the walk is cut at max_steps): what next() hands back
 * -- block_key, the finest node of the cell (depth = block_depth - 1, index; PRUNED or not, like operator[]), x y z =
 * Block::get_point, the node's floats / state / probability; where the block does not exist next() returns false: depth
 * = -1 and x y z = the position tracked so far.  A ray whose start block does not exist has no steps (upstream: n = 0).
 * Upstream's step accounting is kept as written (an xy tie consumes three counts, a step may not move).  Cell indices
 * use 2^(block_depth-1) cells per axis (see la3dm_search about Block::cell_num). */
int la3dm_raycast(la3dm_map *map, const float *start_end, size_t n_rays, size_t max_steps, la3dm_leaf *out,
                  int32_t *n_steps);

/* Inverse of la3dm_export_blocks: fills an EMPTY map from HOST arrays in the reference's Block/OcTree layout
 * (keys[i], nodes[i * nodes_per_block + ...]); afterwards scans can be inserted as if the map had been built here. */
int la3dm_import_blocks(la3dm_map *map, const int64_t *keys, const la3dm_node *nodes, size_t n_blocks);

/* Map serialisation / checkpoint (the reference only has the unused node stream operators,
 * src/bgkoctomap/bgkoctree_node.cpp:46-58).  File = header (magic, method, la3dm_params) + sorted block keys + node
 * arrays as la3dm_export_blocks returns them.  la3dm_load needs an EMPTY map created with the same method and
 * parameters (LA3DM_ERR_INVALID otherwise). */
int la3dm_save(la3dm_map *map, const char *path);
int la3dm_load(la3dm_map *map, const char *path);

/* get_bbox() (src/bgkoctomap/bgkoctomap.cpp:368-381). */
int la3dm_get_bbox(la3dm_map *map, float lim_min[3], float lim_max[3]);

/* Key helpers (src/bgkoctomap/bgkblock.cpp:69-101), evaluated with the map's block size. */
int64_t la3dm_block_to_hash_key(const la3dm_map *map, float x, float y, float z);
void la3dm_hash_key_to_block(const la3dm_map *map, int64_t key, float center[3]);
void la3dm_get_extended_block(const la3dm_map *map, int64_t key, int64_t out7[7]);

/* ---- multi-GPU: test blocks of one scan sharded over ranks ----------------------------------------------------- */
/* Every rank holds a full replica of the map and runs the (cheap) front-end redundantly; rank r predicts the test
 * blocks t with t % world == r.  After insert, each rank packs the node states of ITS blocks; the caller all-gathers
 * the packed buffers (NCCL, e.g. torch.distributed.all_gather_into_tensor on device tensors) and every rank unpacks
 * the peers' rows, so that all replicas are identical again.  Buffers are DEVICE pointers.  pack / unpack only ENQUEUE
 * their kernel on la3dm_stream(map) (no host synchronisation): issue the collective stream-ordered on that stream (or
 * order it with events), and synchronise the stream before reading the map from another stream. */
int la3dm_set_shard(la3dm_map *map, int rank, int world);
int64_t la3dm_shard_row_bytes(const la3dm_map *map);             /* bytes per packed block                 */
int64_t la3dm_shard_rows(const la3dm_map *map);                  /* rows per rank = ceil(T / world)         */
int la3dm_shard_pack(la3dm_map *map, void *d_rows);              /* writes la3dm_shard_rows() rows          */
int la3dm_shard_unpack(la3dm_map *map, const void *d_all_rows);  /* reads world * la3dm_shard_rows() rows   */
void *la3dm_stream(la3dm_map *map);                              /* cudaStream_t the map enqueues on        */
/* Orders the map's stream after everything enqueued so far on producer_stream (a cudaStream_t; NULL = the legacy
 * default stream): event record + stream wait, no host synchronisation. */
int la3dm_stream_wait(la3dm_map *map, void *producer_stream);

/* ---- multi-GPU, BGKOctoMap: replicas kept identical by peer stores from inside the predict kernel ------------------ */
/* Every rank holds a replica and runs the front-end; rank r predicts the test blocks t % world == r and its predict
 * kernel stores every updated node (and the state bytes of every block it changed) into ALL replicas' pools through
 * peer-mapped pointers (NVLink), then flags the scan as complete in every peer's memory; la3dm_insert_pointcloud*
 * returns once all peers have flagged it too.  No pack / collective / unpack, nothing for the caller to do per scan --
 * but every rank MUST insert the same scans in the same order (the replicas, slot numbering included, are a function of
 * the scans alone).  Set-up, once:
 *   1. la3dm_reserve_blocks(map, n): room for n blocks up front -- the pool must not move while peers are attached
 *      (an insert that would have to grow it fails with LA3DM_ERR_NOMEM);
 *   2. same process (one thread per GPU): la3dm_peer_local() on every map; other processes: la3dm_peer_ipc_export(),
 *      hand the two 64-byte handles to the peers by whatever transport the ranks share, la3dm_peer_ipc_open() there;
 *   3. la3dm_peer_attach(map, world, rank, pool_bases, flags): arrays of `world` device pointers, entry [rank] ignored.
 * world <= 8 (one NVSwitch domain). */
#define LA3DM_IPC_HANDLE_BYTES 64
int la3dm_reserve_blocks(la3dm_map *map, size_t blocks);
int la3dm_peer_local(la3dm_map *map, void **pool_base, void **flags);
int la3dm_peer_ipc_export(la3dm_map *map, void *handle_pool, void *handle_flags);
int la3dm_peer_ipc_open(la3dm_map *map, const void *handle_pool, const void *handle_flags, void **pool_base,
                        void **flags);
int la3dm_peer_attach(la3dm_map *map, int world, int rank, void *const *pool_bases, void *const *flags);
int la3dm_peer_detach(la3dm_map *map);
/* Deferred mode (set BEFORE la3dm_peer_attach): nothing crosses NVLink during a scan.  A block is owned by one rank (a
 * fixed function of its key), only the owner updates it and marks it dirty; la3dm_peer_sync() -- collective, every
 * rank calls it -- pushes the dirty blocks to all peers as whole records and returns when all replicas are identical
 * again.  Between an insert and the next sync a replica holds the latest state of ITS blocks only, and the read-side
 * calls (export, search, save, num_leaves) fail with LA3DM_ERR_INVALID.  Use it when most of the map changes every
 * scan (BASELINE.json configs[4]: 3e10 bytes of records per scan): replicating that per scan costs more than the
 * update itself; the default (eager) mode suits scans that touch a small part of the map. */
int la3dm_peer_set_deferred(la3dm_map *map, int deferred);
int la3dm_peer_sync(la3dm_map *map);

/* ---- measurement helper -------------------------------------------------------------------------------------- */
/* FP32 FMA throughput of `device` (TFLOP/s, 2 flop per FMA) from a register-resident FMA loop timed with CUDA
 * events: the denominator for the fp32-pipe roofline of the predict kernel (SURVEY.md section 8d asks for it to be
 * measured on the box).  Not part of the reference's API. */
int la3dm_bench_fp32_peak(int device, float *tflops);

#ifdef __cplusplus
}
#endif
#endif /* LA3DM_B200_H */
