// TEST INFRASTRUCTURE ONLY -- stand-in shadowing include/bgkloctomap/bgklinference.h for the oracle/_ref build.
// Restates train (:32-49), predict (:81-91), point_to_line_dist (:106-141), covSparseLine (:183-197).
#ifndef LA3DM_BGKL_H
#define LA3DM_BGKL_H
#include <cassert>
#include <vector>
#include "standin_math.h"
#include "standin_line.h"

namespace la3dm {
    template<int dim, typename T>
    class BGKLInference {
    public:
        BGKLInference(T sf2, T ell) : sf2(sf2), ell(ell), trained(false) { }

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
            std::vector<T> krow(n);
            for (size_t i = 0; i < m; ++i) {
                point3f p(xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]);
                for (size_t j = 0; j < n; ++j) krow[j] = la3dm_standin::point_to_segment(p, &x[6 * j]);
                for (size_t j = 0; j < n; ++j) {
                    const T d = krow[j] / ell;                                   // Kxz /= ell
                    T k = la3dm_standin::sparse_kernel_unclamped(d, sf2);
                    krow[j] = k < 0.0f ? 0.0f : k;
                }
                T yb = 0.0f, kb = 0.0f;
                for (size_t j = 0; j < n; ++j) { yb += krow[j] * y[j]; kb += krow[j]; }
                ybar[i] = yb; kbar[i] = kb;
            }
        }

    private:
        T sf2, ell;
        std::vector<T> x, y;
        bool trained;
    };

    typedef BGKLInference<3, float> BGKL3f;
}
#endif // LA3DM_BGKL_H
