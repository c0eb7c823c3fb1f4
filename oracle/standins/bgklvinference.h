// TEST INFRASTRUCTURE ONLY -- stand-in shadowing include/bgklvoctomap/bgklvinference.h for the oracle/_ref build.
// Restates train, predict, point_to_line_dist (:98-135) and covSparseLine (:143-157: distance clamped to 1 BEFORE the
// kernel, kernel value never clamped).
#ifndef LA3DM_BGKLV_H
#define LA3DM_BGKLV_H
#include <cassert>
#include <vector>
#include "standin_math.h"
#include "standin_line.h"

namespace la3dm {
    template<int dim, typename T>
    class BGKLVInference {
    public:
        BGKLVInference(T sf2, T ell) : sf2(sf2), ell(ell), trained(false) { }

        void train(const std::vector<T> &x, const std::vector<T> &y) {
            assert(x.size() % (2 * dim) == 0 && (int) (x.size() / (2 * dim)) == (int) y.size());
            this->x = x;
            this->y = y;
            trained = true;
        }

        void predict(const std::vector<T> &xs, std::vector<T> &ybar, std::vector<T> &kbar) const {
            assert(trained == true);
            const size_t m = xs.size() / dim, n = y.size();
            ybar.assign(m, 0.0f);
            kbar.assign(m, 0.0f);
            for (size_t i = 0; i < m; ++i) {
                point3f p(xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]);
                T yb = 0.0f, kb = 0.0f;
                for (size_t j = 0; j < n; ++j) {
                    T d = la3dm_standin::point_to_segment(p, &x[6 * j]) / ell;
                    if (d > 1.0) d = 1.0f;
                    const T k = la3dm_standin::sparse_kernel_unclamped(d, sf2);
                    yb += k * y[j]; kb += k;
                }
                ybar[i] = yb; kbar[i] = kb;
            }
        }

    private:
        T sf2, ell;
        std::vector<T> x, y;
        bool trained;
    };

    typedef BGKLVInference<3, float> BGKLV3f;
}
#endif // LA3DM_BGKLV_H
