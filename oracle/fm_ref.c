/* Plain-C restatement of the reference's FM block (TEST INFRASTRUCTURE — see oracle/__init__.py).
 *
 * Follows FM.forward, /root/reference/models/rank/deepfm/net.py:105-139, statement by statement,
 * with double accumulation, and its analytic backward:
 *   feat[b,f,:]   = W[ids[b,f],:]           (zeros when ids == pad: Embedding(padding_idx=0))
 *   feat[b,F+j,:] = dense[b,j]*dense_w[j,:]                                        net.py:118-120
 *   y1[b] = sum_f W1[ids] + sum_j dense*dense_w1                                   net.py:108-114
 *   y2[b] = 0.5*sum_d ( (sum_n feat)^2 - sum_n feat^2 )                            net.py:123-137
 *   dfeat[b,n,:] = g2[b]*(S[b,:]-feat[b,n,:]) + dfeat_dnn[b,n,:]
 *   dW[ids] += dfeat (not for pad), dW1[ids] += g1, ddense_w[j] += dense*dfeat, ddense_w1[j] += g1*dense
 * Independent of oracle/nets.py (no autograd): used to cross-check the oracle and as a
 * single-thread scalar CPU reference.  Build: make -C oracle  ->  oracle/_build/libfm_ref.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void fm_ref_fwd(const float* W, const float* W1, const int64_t* ids, const float* dense,
                const float* dense_w, const float* dense_w1, float* feat, float* y1, float* y2,
                int64_t B, int F, int Dn, int D, int64_t pad) {
  const int N = F + Dn;
  for (int64_t b = 0; b < B; ++b) {
    double first = 0.0, second = 0.0;
    float* fb = feat + (size_t)b * N * D;
    for (int f = 0; f < F; ++f) {
      const int64_t id = ids[b * F + f];
      for (int d = 0; d < D; ++d) fb[f * D + d] = (id == pad) ? 0.f : W[(size_t)id * D + d];
      if (id != pad) first += W1[id];
    }
    for (int j = 0; j < Dn; ++j) {
      const float x = dense[b * Dn + j];
      for (int d = 0; d < D; ++d) fb[(F + j) * D + d] = x * dense_w[j * D + d];
      first += (double)(x * dense_w1[j]);
    }
    for (int d = 0; d < D; ++d) {
      double s = 0.0, q = 0.0;
      for (int n = 0; n < N; ++n) {
        const double e = fb[n * D + d];
        s += e;
        q += e * e;
      }
      second += s * s - q;
    }
    y1[b] = (float)first;
    y2[b] = (float)(0.5 * second);
  }
}

/* dense gradients: dW [V,D], dW1 [V], ddense_w [Dn,D], ddense_w1 [Dn] must be zero-initialised */
void fm_ref_bwd(const int64_t* ids, const float* dense, const float* feat, const float* dfeat_dnn,
                const float* g1, const float* g2, double* dW, double* dW1, double* ddense_w,
                double* ddense_w1, int64_t B, int F, int Dn, int D, int64_t pad) {
  const int N = F + Dn;
  double* S = (double*)malloc(sizeof(double) * D);
  for (int64_t b = 0; b < B; ++b) {
    const float* fb = feat + (size_t)b * N * D;
    for (int d = 0; d < D; ++d) {
      double s = 0.0;
      for (int n = 0; n < N; ++n) s += fb[n * D + d];
      S[d] = s;
    }
    for (int n = 0; n < N; ++n) {
      for (int d = 0; d < D; ++d) {
        const double df = (double)g2[b] * (S[d] - fb[n * D + d]) +
                          (dfeat_dnn ? (double)dfeat_dnn[((size_t)b * N + n) * D + d] : 0.0);
        if (n < F) {
          const int64_t id = ids[b * F + n];
          if (id != pad) dW[(size_t)id * D + d] += df;
        } else {
          ddense_w[(n - F) * D + d] += (double)dense[b * Dn + n - F] * df;
        }
      }
      if (n < F) {
        const int64_t id = ids[b * F + n];
        if (id != pad) dW1[id] += (double)g1[b];
      } else {
        ddense_w1[n - F] += (double)g1[b] * dense[b * Dn + n - F];
      }
    }
  }
  free(S);
}
