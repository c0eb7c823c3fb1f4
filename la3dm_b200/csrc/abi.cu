// la3dm_b200 -- extern "C" boundary (include/la3dm_b200.h).  Thin: argument checks, exception -> status code.
#include <cmath>
#include <cstring>
#include <new>

#include "engine.cuh"

using la3dm_b200::Map;


namespace {

template <typename F>
int guarded(la3dm_map *map, F &&f) {
    try {
        f();
        return LA3DM_OK;
    } catch (const la3dm_b200::CudaError &e) {
        if (map) {
            char buf[512];
            snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d", (int) e.code, cudaGetErrorString(e.code), e.file,
                     e.line);
            map->m.last_error = buf;
        }
        cudaGetLastError();
        return LA3DM_ERR_CUDA;
    } catch (const la3dm_b200::StatusError &e) {
        if (map) map->m.last_error = e.msg;
        return e.status;
    } catch (const std::bad_alloc &) {
        if (map) map->m.last_error = "host allocation failed";
        return LA3DM_ERR_NOMEM;
    } catch (...) {
        if (map) map->m.last_error = "unknown failure";
        return LA3DM_ERR_INVALID;
    }
}

thread_local std::string g_create_error;

}  // namespace

extern "C" {

int la3dm_abi_version(void) { return LA3DM_B200_ABI_VERSION; }

const char *la3dm_status_string(int s) {
    switch (s) {
        case LA3DM_OK: return "ok";
        case LA3DM_ERR_INVALID: return "invalid argument";
        case LA3DM_ERR_CUDA: return "CUDA failure";
        case LA3DM_ERR_UNSUPPORTED: return "unsupported";
        case LA3DM_ERR_EXTENT: return "scan extent too large";
        case LA3DM_ERR_NOMEM: return "out of memory";
        case LA3DM_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

int la3dm_create(int method, const la3dm_params *params, int device, la3dm_map **out) {
    if (!params || !out) return LA3DM_ERR_INVALID;
    *out = nullptr;
    la3dm_map *map = new (std::nothrow) la3dm_map();
    if (!map) return LA3DM_ERR_NOMEM;
    const int rc = guarded(map, [&] { map->m.init(method, *params, device); });
    if (rc != LA3DM_OK) {
        g_create_error = map->m.last_error;
        delete map;
        return rc;
    }
    *out = map;
    return LA3DM_OK;
}

int la3dm_destroy(la3dm_map *map) {
    if (!map) return LA3DM_ERR_INVALID;
    cudaSetDevice(map->m.device);
    if (map->m.peer_wait_pending) {          // peers may still be storing into this replica's pool
        try { map->m.peer_wait_now(); } catch (...) { cudaGetLastError(); }
    }
    if (map->m.stream) cudaStreamSynchronize(map->m.stream);
    delete map;
    return LA3DM_OK;
}

const char *la3dm_last_error(const la3dm_map *map) { return map ? map->m.last_error.c_str() : g_create_error.c_str(); }

int la3dm_insert_pointcloud_device(la3dm_map *map, const float *d_xyz, size_t n, size_t stride_bytes,
                                   const float origin[3], float ds_resolution, float free_res, float max_range) {
    if (!map || !origin) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        map->m.insert_device(d_xyz, n, stride_bytes, origin, ds_resolution, free_res, max_range, 0);
    });
}

int la3dm_insert_pointcloud(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float origin[3],
                            float ds_resolution, float free_res, float max_range) {
    if (!map || !origin || (n && !xyz)) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        LA3DM_CUDA(cudaSetDevice(m.device));
        if (stride_bytes < 12 || stride_bytes % 4) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "bad stride_bytes"};
        m.cloud.reserve(n * stride_bytes + 16, m.stream);
        if (n) LA3DM_CUDA(cudaMemcpyAsync(m.cloud.p, xyz, n * stride_bytes, cudaMemcpyHostToDevice, m.stream));
        m.h2d_bytes = (long long) (n * stride_bytes);
        m.insert_device(m.cloud.as<float>(), n, stride_bytes, origin, ds_resolution, free_res, max_range, 0);
    });
}

int la3dm_insert_pointcloud_ingest(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float tf[12],
                                   float prefilter_ds, int min_points, const float origin[3], float ds_resolution,
                                   float free_res, float max_range) {
    if (!map || !origin || !tf || (n && !xyz)) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        LA3DM_CUDA(cudaSetDevice(m.device));
        if (stride_bytes < 12 || stride_bytes % 4) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "bad stride_bytes"};
        m.cloud.reserve(n * stride_bytes + 16, m.stream);
        if (n) LA3DM_CUDA(cudaMemcpyAsync(m.cloud.p, xyz, n * stride_bytes, cudaMemcpyHostToDevice, m.stream));
        m.h2d_bytes = (long long) (n * stride_bytes);
        if (m.ingest_pre_ds != prefilter_ds) m.invalidate_graph();      // with / without the prefilter: other kernels
        memcpy(m.ingest_tf, tf, sizeof(m.ingest_tf));
        m.ingest_pre_ds = prefilter_ds;
        m.ingest_min_points = min_points;
        m.insert_device(m.cloud.as<float>(), n, stride_bytes, origin, ds_resolution, free_res, max_range, 3);
    });
}

int la3dm_insert_training_data(la3dm_map *map, const float *xyzy, size_t n, size_t stride_bytes) {
    if (!map || (n && !xyzy)) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        LA3DM_CUDA(cudaSetDevice(m.device));
        if (stride_bytes < 16 || stride_bytes % 4) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "bad stride_bytes"};
        m.cloud.reserve(n * stride_bytes + 16, m.stream);
        if (n) LA3DM_CUDA(cudaMemcpyAsync(m.cloud.p, xyzy, n * stride_bytes, cudaMemcpyHostToDevice, m.stream));
        m.h2d_bytes = (long long) (n * stride_bytes);
        const float o[3] = {0.f, 0.f, 0.f};
        m.insert_device(m.cloud.as<float>(), n, stride_bytes, o, -1.f, 1.f, -1.f, 2);
    });
}

int la3dm_insert_training_data_device(la3dm_map *map, const float *d_xyzy, size_t n, size_t stride_bytes) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        const float o[3] = {0.f, 0.f, 0.f};
        map->m.insert_device(d_xyzy, n, stride_bytes, o, -1.f, 1.f, -1.f, 2);
    });
}

int la3dm_training_data(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, const float origin[3],
                        float ds_resolution, float free_res, float max_range, float *out, size_t capacity,
                        size_t *n_out) {
    if (!map || !origin || (n && !xyz)) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        LA3DM_CUDA(cudaSetDevice(m.device));
        if (stride_bytes < 12 || stride_bytes % 4) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "bad stride_bytes"};
        m.cloud.reserve(n * stride_bytes + 16, m.stream);
        if (n) LA3DM_CUDA(cudaMemcpyAsync(m.cloud.p, xyz, n * stride_bytes, cudaMemcpyHostToDevice, m.stream));
        m.insert_device(m.cloud.as<float>(), n, stride_bytes, origin, ds_resolution, free_res, max_range, 1);
        const size_t nt = (size_t) m.stats.n_train;
        if (n_out) *n_out = nt;
        if (!out || nt == 0) return;
        if (capacity < nt) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "training_data: capacity too small"};
        std::vector<float4> h(nt);
        LA3DM_CUDA(cudaMemcpy(h.data(), m.xy.p, nt * sizeof(float4), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nt; ++i) {
            float *o = out + 7 * i;
            o[0] = o[3] = h[i].x; o[1] = o[4] = h[i].y; o[2] = o[5] = h[i].z; o[6] = h[i].w;
        }
    });
}

int la3dm_training_rays(la3dm_map *map, float *rays, size_t ray_capacity, size_t *n_rays, int32_t *ray_idx,
                        size_t idx_capacity) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        if (m.hp.method != LA3DM_BGKL && m.hp.method != LA3DM_BGKLV)
            throw la3dm_b200::StatusError{LA3DM_ERR_UNSUPPORTED, "training_rays: BGKL / BGKLV only"};
        LA3DM_CUDA(cudaSetDevice(m.device));
        LA3DM_CUDA(cudaStreamSynchronize(m.stream));
        const size_t nr = m.hp.method == LA3DM_BGKL ? (size_t) m.h_cnt->n_hits : (size_t) m.h_cnt->n_frees;
        const size_t nt = (size_t) m.h_cnt->n_train;
        if (n_rays) *n_rays = nr;
        if (rays && nr) {
            if (ray_capacity < nr) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "training_rays: capacity too small"};
            std::vector<float4> h(2 * nr);
            LA3DM_CUDA(cudaMemcpy(h.data(), m.rays.p, 2 * nr * sizeof(float4), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nr; ++i) {
                float *o = rays + 6 * i;
                o[0] = h[2 * i].x; o[1] = h[2 * i].y; o[2] = h[2 * i].z;
                o[3] = h[2 * i + 1].x; o[4] = h[2 * i + 1].y; o[5] = h[2 * i + 1].z;
            }
        }
        if (ray_idx && nt) {
            if (idx_capacity < nt) throw la3dm_b200::StatusError{LA3DM_ERR_INVALID, "training_rays: capacity too small"};
            LA3DM_CUDA(cudaMemcpy(ray_idx, m.ray_of.p, nt * sizeof(int32_t), cudaMemcpyDeviceToHost));
        }
    });
}

int la3dm_last_stats(const la3dm_map *map, la3dm_scan_stats *out) {
    if (!map || !out) return LA3DM_ERR_INVALID;
    *out = map->m.stats;
    return LA3DM_OK;
}

int64_t la3dm_num_blocks(const la3dm_map *map) { return map ? map->m.n_blocks : -1; }
int32_t la3dm_nodes_per_block(const la3dm_map *map) { return map ? map->m.hp.nodes : -1; }

int la3dm_export_blocks(la3dm_map *map, int64_t *keys, la3dm_node *nodes, size_t capacity_blocks, size_t *n_blocks) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.export_blocks(keys, nodes, capacity_blocks, n_blocks); });
}

int64_t la3dm_num_leaves(la3dm_map *map) {
    if (!map) return -1;
    long long n = -1;
    const int rc = guarded(map, [&] { n = map->m.count_leaves(); });
    return rc == LA3DM_OK ? n : rc;
}

int la3dm_export_leaves(la3dm_map *map, la3dm_leaf *out, size_t capacity, size_t *n_out) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.export_leaves(out, capacity, n_out); });
}

int la3dm_export_touched(la3dm_map *map, unsigned int state_mask, la3dm_leaf *leaves, size_t capacity_leaves,
                         size_t *n_leaves, int64_t *block_keys, size_t capacity_blocks, size_t *n_blocks, int clear) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        map->m.export_touched(state_mask, leaves, capacity_leaves, n_leaves, block_keys, capacity_blocks, n_blocks, clear != 0);
    });
}

int la3dm_search(la3dm_map *map, const float *xyz, size_t n, size_t stride_bytes, int finest_only, la3dm_leaf *out) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.search(xyz, n, stride_bytes, false, finest_only, out); });
}

int la3dm_raycast(la3dm_map *map, const float *start_end, size_t n_rays, size_t max_steps, la3dm_leaf *out,
                  int32_t *n_steps) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.raycast(start_end, n_rays, max_steps, out, n_steps); });
}

int la3dm_import_blocks(la3dm_map *map, const int64_t *keys, const la3dm_node *nodes, size_t n_blocks) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.import_blocks(keys, nodes, n_blocks); });
}

int la3dm_save(la3dm_map *map, const char *path) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.save(path); });
}

int la3dm_load(la3dm_map *map, const char *path) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] { map->m.load(path); });
}

int64_t la3dm_block_to_hash_key(const la3dm_map *map, float x, float y, float z) {
    if (!map) return -1;
    const float bs = map->m.hp.block_size;
    return la3dm_b200::make_key(la3dm_b200::axis_index(x, bs), la3dm_b200::axis_index(y, bs),
                                la3dm_b200::axis_index(z, bs));
}

void la3dm_hash_key_to_block(const la3dm_map *map, int64_t key, float c[3]) {
    if (!map || !c) return;
    const float bs = map->m.hp.block_size;
    c[0] = la3dm_b200::axis_center(key >> 40, bs);
    c[1] = la3dm_b200::axis_center((key >> 20) & 0xFFFFF, bs);
    c[2] = la3dm_b200::axis_center(key & 0xFFFFF, bs);
}

// get_extended_block (src/bgkoctomap/bgkblock.cpp:85-101)
void la3dm_get_extended_block(const la3dm_map *map, int64_t key, int64_t out7[7]) {
    if (!map || !out7) return;
    float c[3];
    la3dm_hash_key_to_block(map, key, c);
    const float bs = map->m.hp.block_size;
    out7[0] = key;
    for (int i = 0; i < 6; ++i) {
        const float s = (i % 2 == 0) ? bs : -bs;
        const float ex = (i / 2 == 0) ? s : 0.f, ey = (i / 2 == 1) ? s : 0.f, ez = (i / 2 == 2) ? s : 0.f;
        out7[i + 1] = la3dm_block_to_hash_key(map, ex + c[0], ey + c[1], ez + c[2]);
    }
}

// get_bbox (src/bgkoctomap/bgkoctomap.cpp:368-381): min/max of block centres -/+ block_size * 0.5
int la3dm_get_bbox(la3dm_map *map, float lim_min[3], float lim_max[3]) {
    if (!map || !lim_min || !lim_max) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        for (int a = 0; a < 3; ++a) lim_min[a] = lim_max[a] = 0.f;
        const size_t n = (size_t) m.n_blocks;
        if (n == 0) return;
        std::vector<int64_t> keys(n);
        LA3DM_CUDA(cudaSetDevice(m.device));
        LA3DM_CUDA(cudaStreamSynchronize(m.stream));
        LA3DM_CUDA(cudaMemcpy(keys.data(), m.keys.p, n * 8, cudaMemcpyDeviceToHost));
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (size_t i = 0; i < n; ++i) {
            float c[3];
            la3dm_hash_key_to_block(map, keys[i], c);
            for (int a = 0; a < 3; ++a) { mn[a] = std::fmin(mn[a], c[a]); mx[a] = std::fmax(mx[a], c[a]); }
        }
        const float bs = m.hp.block_size;
        for (int a = 0; a < 3; ++a) {
            // point3f(bs,bs,bs) * 0.5 : operator*(float) on each component, then -= / +=
            const float h = bs * 0.5f;
            lim_min[a] = mn[a] - h;
            lim_max[a] = mx[a] + h;
        }
    });
}

int la3dm_set_shard(la3dm_map *map, int rank, int world) {
    if (!map || world < 1 || rank < 0 || rank >= world) return LA3DM_ERR_INVALID;
    // BGKLV predicts per voxel through its own block list (k_lv_blocks / k_lv_predict), which does not honour the shard
    // and never fills the neighbour plan the exchange indexes: refuse instead of exchanging garbage
    if (world > 1 && map->m.hp.method == LA3DM_BGKLV) {
        map->m.last_error = "sharding is not implemented for BGKLV";
        return LA3DM_ERR_UNSUPPORTED;
    }
    map->m.shard_rank = rank;
    map->m.shard_world = world;
    return LA3DM_OK;
}

void *la3dm_stream(la3dm_map *map) { return map ? (void *) map->m.stream : nullptr; }

int la3dm_stream_wait(la3dm_map *map, void *producer_stream) {
    if (!map) return LA3DM_ERR_INVALID;
    return guarded(map, [&] {
        Map &m = map->m;
        LA3DM_CUDA(cudaSetDevice(m.device));
        LA3DM_CUDA(cudaEventRecord(m.ev_wait, (cudaStream_t) producer_stream));
        LA3DM_CUDA(cudaStreamWaitEvent(m.stream, m.ev_wait, 0));
    });
}

}  // extern "C"

static_assert(sizeof(la3dm_node) == 16, "la3dm_node must match the reference's 16-byte Occupancy");
static_assert(sizeof(la3dm_leaf) == 56, "la3dm_leaf layout");
