// TEST INFRASTRUCTURE (CPU): arithmetic claims the fused front-end (la3dm_b200/csrc/frontend_fused.cu) rests on, checked
// against plain sequential fp32 evaluation on the host.  Built by tests/test_host_logic.py with -ffp-contract=off.
//   1. add_repeat(s, x, n) == s + x + x + ... (n additions, one rounding each)                     (common.cuh)
//   2. the sensor origin's voxel: n_hits copies of the origin interleaved with a few other samples, summed in push order,
//      == the reconstruction from the other samples alone (each knows how many origins precede it)      (vg_centroid<1>)
//   3. the bounding box of a beam's free points == the box of {origin, nearest sample, farthest regular sample, tail
//      sample}: every sample is fl(o + fl(n * d)) and rounding is monotonic in d                        (centroid_done<0>)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../la3dm_b200/csrc/common.cuh"

using la3dm_b200::add_repeat;

static float seq_add(float s, float x, unsigned int n) {
    volatile float a = s;
    for (unsigned int i = 0; i < n; ++i) a = a + x;
    return a;
}

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long bad = 0, checks = 0;
    // ---- 1
    for (int it = 0; it < 20000; ++it) {
        const double mag = std::pow(10.0, -6.0 + 12.0 * U(rng));
        float x = (float) ((U(rng) < 0.5 ? -1.0 : 1.0) * mag * (0.5 + U(rng)));
        float s = (float) ((U(rng) < 0.3 ? -1.0 : 1.0) * std::pow(10.0, -6.0 + 12.0 * U(rng)) * U(rng));
        if (it % 7 == 0) s = 0.f;
        if (it % 11 == 0) x = (float) (0.05 + 50.0 * U(rng));           // coordinates of a sensor origin
        const unsigned int n = (unsigned int) (1 + (it % 5 == 0 ? 200000 : 3000) * U(rng));
        const float a = add_repeat(s, x, n), b = seq_add(s, x, n);
        ++checks;
        if (la3dm_b200::f2u(a) != la3dm_b200::f2u(b)) { if (++bad < 5) printf("add_repeat(%.9g, %.9g, %u) = %.9g, sequential %.9g\n", s, x, n, a, b); }
    }
    // ---- 2
    for (int it = 0; it < 2000; ++it) {
        const unsigned int n_hits = 1 + (unsigned int) (60000 * U(rng));
        const float o = (float) (-50.0 + 100.0 * U(rng));
        const int n_other = (int) (6 * U(rng));
        // other samples: (ordinal of their hit, value); several may belong to the same hit, ordinals ascending
        std::vector<std::pair<unsigned int, float>> others;
        unsigned int h = 0;
        for (int k = 0; k < n_other; ++k) {
            h += (unsigned int) (U(rng) * n_hits / (n_other + 1));
            if (h >= n_hits) h = n_hits - 1;
            others.emplace_back(h, o + (float) (0.1 * (U(rng) - 0.5)));
        }
        // push order: for every kept hit its origin copy, then its samples
        volatile float ref = 0.f;
        size_t q = 0;
        for (unsigned int hh = 0; hh < n_hits; ++hh) {
            ref = ref + o;
            while (q < others.size() && others[q].first == hh) { ref = ref + others[q].second; ++q; }
        }
        float acc = 0.f;
        unsigned int done = 0;
        for (auto &ov : others) {
            const unsigned int k = ov.first + 1u;                       // origins pushed before this sample
            acc = add_repeat(acc, o, k - done);
            { volatile float t = acc; t = t + ov.second; acc = t; }
            done = k;
        }
        acc = add_repeat(acc, o, n_hits - done);
        ++checks;
        const float r = ref;
        if (la3dm_b200::f2u(acc) != la3dm_b200::f2u(r)) { if (++bad < 10) printf("origin voxel: %.9g vs sequential %.9g (n_hits %u, %d others)\n", acc, r, n_hits, n_other); }
    }
    // ---- 3
    for (int it = 0; it < 20000; ++it) {
        const float fr = (float) (0.03 + 0.6 * U(rng));
        float o[3], hit[3], nrm[3];
        for (int a = 0; a < 3; ++a) { o[a] = (float) (-30 + 60 * U(rng)); hit[a] = o[a] + (float) ((U(rng) - 0.5) * (it % 3 ? 60.0 : 2.0)); }
        const float dx = hit[0] - o[0], dy = hit[1] - o[1], dz = hit[2] - o[2];
        const float l = (float) std::sqrt((double) (dx * dx + dy * dy + dz * dz));
        if (!(l > 0)) continue;
        nrm[0] = dx / l; nrm[1] = dy / l; nrm[2] = dz / l;
        // beam_sample (bgkoctomap.cpp:433-458): d = fr; while (d < l) { push o + n * d; d += fr; }  then the tail at l - fr
        std::vector<float> ds;
        { volatile float d = fr; while (d < l && ds.size() < 100000) { ds.push_back((float) d); d = d + fr; } }
        const size_t n_reg = ds.size();
        if (l > fr) ds.push_back(l - fr);
        float mn[3], mx[3], cmn[3], cmx[3];
        for (int a = 0; a < 3; ++a) { mn[a] = mx[a] = cmn[a] = cmx[a] = o[a]; }           // the origin copy
        auto sample = [&](float d, int a) { volatile float p = nrm[a] * d; volatile float s = o[a] + p; return (float) s; };
        for (float d : ds) for (int a = 0; a < 3; ++a) { const float s = sample(d, a); mn[a] = std::fmin(mn[a], s); mx[a] = std::fmax(mx[a], s); }
        std::vector<float> cand;
        if (n_reg) { cand.push_back(ds[0]); cand.push_back(ds[n_reg - 1]); }
        if (l > fr) cand.push_back(l - fr);
        for (float d : cand) for (int a = 0; a < 3; ++a) { const float s = sample(d, a); cmn[a] = std::fmin(cmn[a], s); cmx[a] = std::fmax(cmx[a], s); }
        ++checks;
        for (int a = 0; a < 3; ++a)
            if (mn[a] != cmn[a] || mx[a] != cmx[a]) { if (++bad < 15) printf("beam box axis %d: [%.9g, %.9g] vs candidates [%.9g, %.9g] (l %.9g fr %.9g)\n", a, mn[a], mx[a], cmn[a], cmx[a], l, fr); break; }
    }
    printf("%ld checks, %ld bad\n", checks, bad);
    return bad ? 1 : 0;
}
