// TEST INFRASTRUCTURE ONLY -- stand-in shadowing include/gpoctomap/gpregressor.h for the oracle/_ref build.
// Restates GPRegressor::train (:42-51: K = Matern-3/2 + noise*I, LLT, alpha = K^-1 y, keep L), predict (:80-92:
// Ks = k(x, xs) [N x M]; m = Ks^T alpha; v = L^-1 Ks; var = sf2 - diag(v^T v)) and covMaterniso3 (:114-117:
// coordinates scaled by float(1.73205/ell) first, k = (1 + r) * exp(-r) * sf2) as plain fp32 loops with a
// left-looking (dot-product form) Cholesky; every dot product is summed sequentially in index order.
#ifndef LA3DM_GP_REGRESSOR_H
#define LA3DM_GP_REGRESSOR_H
#include <cassert>
#include <cmath>
#include <vector>
#include "standin_math.h"

namespace la3dm {
    template<int dim, typename T>
    class GPRegressor {
    public:
        GPRegressor(T sf2, T ell, T noise) : sf2(sf2), ell(ell), noise(noise), trained(false) { }

        void train(const std::vector<T> &x, const std::vector<T> &y) {
            assert(x.size() % dim == 0 && (int) (x.size() / dim) == (int) y.size());
            n = y.size();
            const T scale = (T) (1.73205 / ell);
            xs_.resize(3 * n);
            for (size_t i = 0; i < 3 * n; ++i) xs_[i] = scale * x[i];
            L.assign(n * n, 0.0f);
            // K + noise*I (lower triangle is all we need)
            for (size_t i = 0; i < n; ++i)
                for (size_t j = 0; j <= i; ++j) {
                    T k = kern(&xs_[3 * i], &xs_[3 * j]);
                    if (i == j) k = k + noise * 1.0f;
                    L[i * n + j] = k;
                }
            // Cholesky, row by row: L_ij = (K_ij - sum_k<j L_ik L_jk) / L_jj ; L_ii = sqrt(K_ii - sum_k<i L_ik^2)
            for (size_t i = 0; i < n; ++i) {
                for (size_t j = 0; j <= i; ++j) {
                    T s = L[i * n + j];
                    for (size_t k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
                    L[i * n + j] = (i == j) ? std::sqrt(s) : s / L[j * n + j];
                }
            }
            // alpha = L^-T (L^-1 y)
            alpha.assign(n, 0.0f);
            for (size_t i = 0; i < n; ++i) {
                T s = y[i];
                for (size_t k = 0; k < i; ++k) s -= L[i * n + k] * alpha[k];
                alpha[i] = s / L[i * n + i];
            }
            for (size_t ii = n; ii-- > 0;) {
                T s = alpha[ii];
                for (size_t k = ii + 1; k < n; ++k) s -= L[k * n + ii] * alpha[k];
                alpha[ii] = s / L[ii * n + ii];
            }
            trained = true;
        }

        void predict(const std::vector<T> &xs, std::vector<T> &m, std::vector<T> &var) const {
            assert(trained == true);
            const size_t M = xs.size() / dim;
            m.assign(M, 0.0f);
            var.assign(M, 0.0f);
            const T scale = (T) (1.73205 / ell);
            std::vector<T> ks(n), v(n);
            for (size_t c = 0; c < M; ++c) {
                const T q[3] = {scale * xs[3 * c], scale * xs[3 * c + 1], scale * xs[3 * c + 2]};
                for (size_t i = 0; i < n; ++i) ks[i] = kern(&xs_[3 * i], q);
                T mu = 0.0f;
                for (size_t i = 0; i < n; ++i) mu += ks[i] * alpha[i];
                T vv = 0.0f;
                for (size_t i = 0; i < n; ++i) {
                    T s = ks[i];
                    for (size_t k = 0; k < i; ++k) s -= L[i * n + k] * v[k];
                    v[i] = s / L[i * n + i];
                    vv += v[i] * v[i];
                }
                m[c] = mu;
                var[c] = sf2 - vv;
            }
        }

    private:
        // covMaterniso3 element: r = || z' - x' ||, (1 + r) * exp(-r) * sf2   (dist(): rows of z minus row of x)
        T kern(const T *a, const T *b) const {
            const T r = la3dm_standin::norm3(b[0] - a[0], b[1] - a[1], b[2] - a[2]);
            return ((1 + r) * la3dm_standin::exp_f(-r)) * sf2;
        }

        T sf2, ell, noise;
        size_t n = 0;
        std::vector<T> xs_, L, alpha;
        bool trained;
    };

    typedef GPRegressor<3, float> GPR3f;
}
#endif // LA3DM_GP_REGRESSOR_H
