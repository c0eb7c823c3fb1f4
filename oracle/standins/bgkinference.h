// TEST INFRASTRUCTURE ONLY -- stand-in that shadows the reference's include/bgkoctomap/bgkinference.h when building
// oracle/_ref (the reference .cpp includes it by bare name, so an earlier -I wins).  Same class name, constructor and
// train/predict signatures; the Eigen expressions of bgkinference.h:28-44 (train), :73-79 (predict), :88-93 (dist),
// :113-126 (covSparse) are restated as plain fp32 loops.  See standin_math.h for the two arithmetic flavours.
#ifndef LA3DM_BGK_H
#define LA3DM_BGK_H
#include <cassert>
#include <vector>
#include "standin_math.h"

namespace la3dm {
    template<int dim, typename T>
    class BGKInference {
    public:
        BGKInference(T sf2, T ell) : sf2(sf2), ell(ell), trained(false) { }

        void train(const std::vector<T> &x, const std::vector<T> &y) {
            assert(x.size() % dim == 0 && (int) (x.size() / dim) == (int) y.size());
            const size_t n = y.size();
            // covSparse divides the stored training matrix by ell every call (x / ell, element-wise);
            // doing it once here yields the same fp32 values.
            xs_[0].resize(n); xs_[1].resize(n); xs_[2].resize(n);
            for (size_t j = 0; j < n; ++j)
                for (int k = 0; k < 3; ++k) xs_[k][j] = x[3 * j + k] / ell;
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
                const T px = xs[3 * i] / ell, py = xs[3 * i + 1] / ell, pz = xs[3 * i + 2] / ell;
                const T *zx = xs_[0].data(), *zy = xs_[1].data(), *zz = xs_[2].data();
                T *kr = krow.data();
                for (size_t j = 0; j < n; ++j) {
                    const T d = la3dm_standin::norm3(zx[j] - px, zy[j] - py, zz[j] - pz);
                    T k = la3dm_standin::sparse_kernel_unclamped(d, sf2);
                    kr[j] = k < 0.0f ? 0.0f : k;     // bgkinference.h:120-125
                }
                T yb = 0.0f, kb = 0.0f;              // ybar = Ks*y ; kbar = Ks.rowwise().sum()  (sequential fp32)
                for (size_t j = 0; j < n; ++j) { yb += kr[j] * y[j]; kb += kr[j]; }
                ybar[i] = yb; kbar[i] = kb;
            }
        }

    private:
        T sf2, ell;
        std::vector<T> xs_[3];
        std::vector<T> y;
        bool trained;
    };

    typedef BGKInference<3, float> BGK3f;
}
#endif // LA3DM_BGK_H
