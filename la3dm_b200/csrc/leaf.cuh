// la3dm_b200 -- read side shared by the leaf export and the point query: is_leaf and one leaf as LeafIterator / search
// expose it (get_loc, get_size, get_prob, get_var).
#pragma once
#include "common.cuh"

namespace la3dm_b200 {

// is_leaf (src/bgkoctomap/bgkoctree.cpp:72-82) on the packed state bytes
__device__ inline bool node_is_leaf(const unsigned char *bst, const DevParams &P, int d, int i) {
    if ((bst[P.layer_off[d] + i] & 7) == P.pruned_state) return false;
    if (d + 1 < P.depth) return (bst[P.layer_off[d + 1] + 8 * i] & 7) == P.pruned_state;
    return true;
}

// Occupancy::get_prob / get_var of the three node types from the two stored floats
__device__ inline void node_prob_var(const DevParams &P, float2 v, float &prob, float &var) {
    if (P.method == LA3DM_GP) {
        // gpoctree_node.cpp:31-34, gpoctree_node.h:60
        prob = 1.0f / (1.0f + (float) exp((double) (-P.l * v.x / P.max_ivar)));
        var = 1.0f / v.y;
    } else if (P.method == LA3DM_BGKLV) {
        // bgklvoctree_node.cpp:29-62
        const float W = (v.x + v.y < P.min_W) ? P.min_W : v.x + v.y;
        float pr;
        if (v.x > v.y) pr = (float) ((double) (v.x / (W - v.y)) + (double) (W - v.x - v.y) * 0.5 / (double) (W - v.y));
        else pr = (float) (0.5 * (double) (W - v.y - v.x) / (double) (W - v.x));
        prob = pr;
        var = (float) ((double) (v.x / W) * pow((double) (1 - pr), 2.0) +
                       (double) ((W - v.x - v.y) / W) * pow(0.5 - (double) pr, 2.0) +
                       (double) (v.y / W) * pow((double) pr, 2.0));
    } else {
        prob = v.x / (v.x + v.y);                                        // bgkoctree_node.cpp:27-29
        var = (v.x * v.y) / ((v.x + v.y) * (v.x + v.y) * (v.x + v.y + 1.0f));   // bgkoctree_node.h:60
    }
}

// node (d, i) of the block `key` centred at (cx, cy, cz): v = its two floats, s = its state byte, o = its LUT offset
__device__ inline la3dm_leaf make_leaf(const DevParams &P, long long key, int d, int i, float2 v, unsigned char s,
                                       float3 o, float cx, float cy, float cz) {
    la3dm_leaf L;
    L.block_key = key; L.depth = d; L.index = i;
    L.x = o.x + cx; L.y = o.y + cy; L.z = o.z + cz;                       // Block::get_loc
    L.size = (float) ((double) P.block_size / pow(2.0, (double) d));      // Block::get_size
    L.a = v.x; L.b = v.y;
    node_prob_var(P, v, L.prob, L.var);
    L.state = s & 7; L.classified = s >> 7;
    for (int q = 0; q < 6; ++q) L._pad[q] = 0;
    return L;
}

}  // namespace la3dm_b200
