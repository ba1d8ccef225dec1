/*
 * CPU oracle, C/OpenMP restatement of tatva's element-level hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked by nothing under tatva_b200/; used by tests/ (larger parity
 * sizes than the NumPy oracle can finish in seconds) and by bench.py's cpu_baseline / --impl
 * reference legs.  It follows the reference's arithmetic line by line (paths relative to the
 * tatva v0.11.1 tree):
 *   dNdr(xi)                     tatva/element/base.py:262-265 (Tri3), :467-472 (Tet4), :531-568 (Hex8)
 *   J = dNdr @ X_e, det J        tatva/element/base.py:90-93
 *   dNdX = inv(J) @ dNdr         tatva/element/base.py:111-113
 *   grad = einsum(dn,n...->...d) tatva/element/base.py:114
 *   W = det J * w_q              tatva/operator.py:172-192
 *   E = sum_e sum_q W psi        tatva/operator.py:307-356
 *   gather u[elements]           tatva/operator.py:221 ; its transpose = scatter-add (jax.grad)
 *   psi (linear elastic)         tests/test_sparse.py:20-38
 *   psi (neo-Hookean)            tests/test_sparse_tracer.py:103-115
 * It is checked against the NumPy oracle (itself pinned to reference outputs) in
 * tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { TRI3 = 0, TET4 = 1, HEX8 = 2 };
enum { LINEAR_ELASTIC = 0, NEO_HOOKEAN = 1 };
enum { MODE_ENERGY = 0, MODE_RESIDUAL = 1, MODE_HVP = 2 };

static const double HEX_S[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                   {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};

static void elem_info(int kind, int* dim, int* npe, int* nq) {
  if (kind == TRI3) { *dim = 2; *npe = 3; *nq = 1; }
  else if (kind == TET4) { *dim = 3; *npe = 4; *nq = 1; }
  else { *dim = 3; *npe = 8; *nq = 8; }
}

static double quad_weight(int kind) { return kind == TRI3 ? 0.5 : kind == TET4 ? 1.0 / 6 : 1.0; }

/* dNdr[d*npe + n] at quadrature point q */
static void shape_dfn(int kind, int q, double* dNdr) {
  if (kind == TRI3) {
    const double t[2][3] = {{-1, 1, 0}, {-1, 0, 1}};
    memcpy(dNdr, t, sizeof t);
  } else if (kind == TET4) {
    const double t[3][4] = {{-1, 1, 0, 0}, {-1, 0, 1, 0}, {-1, 0, 0, 1}};
    memcpy(dNdr, t, sizeof t);
  } else {
    const double a = 1.0 / sqrt(3.0);
    const double x = a * HEX_S[q][0], y = a * HEX_S[q][1], z = a * HEX_S[q][2];
    for (int n = 0; n < 8; ++n) {
      const double fx = 1 + HEX_S[n][0] * x, fy = 1 + HEX_S[n][1] * y, fz = 1 + HEX_S[n][2] * z;
      dNdr[0 * 8 + n] = 0.125 * HEX_S[n][0] * fy * fz;
      dNdr[1 * 8 + n] = 0.125 * HEX_S[n][1] * fx * fz;
      dNdr[2 * 8 + n] = 0.125 * HEX_S[n][2] * fx * fy;
    }
  }
}

static double inv_small(int d, const double* A, double* Ai) {
  if (d == 2) {
    const double det = A[0] * A[3] - A[1] * A[2];
    Ai[0] = A[3] / det; Ai[1] = -A[1] / det; Ai[2] = -A[2] / det; Ai[3] = A[0] / det;
    return det;
  }
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  Ai[0] = c00 / det; Ai[3] = c01 / det; Ai[6] = c02 / det;
  Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det;
  Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det;
  Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det;
  Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
  Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
  return det;
}

/* dNdX[c*npe + n], returns det J */
static double geometry(int kind, int dim, int npe, int q, const double* X /* npe x dim */, double* dNdX) {
  double dNdr[24], J[9], Ji[9];
  shape_dfn(kind, q, dNdr);
  for (int d = 0; d < dim; ++d)
    for (int c = 0; c < dim; ++c) {
      double s = 0;
      for (int n = 0; n < npe; ++n) s += dNdr[d * npe + n] * X[n * dim + c];
      J[d * dim + c] = s;
    }
  const double det = inv_small(dim, J, Ji);
  for (int c = 0; c < dim; ++c)
    for (int n = 0; n < npe; ++n) {
      double s = 0;
      for (int d = 0; d < dim; ++d) s += Ji[c * dim + d] * dNdr[d * npe + n];
      dNdX[c * npe + n] = s;
    }
  return det;
}

static double psi_le(int d, const double* G, double mu, double lm) {
  double tr = 0, ee = 0;
  for (int i = 0; i < d; ++i) {
    tr += G[i * d + i];
    for (int j = 0; j < d; ++j) {
      const double e = 0.5 * (G[i * d + j] + G[j * d + i]);
      ee += e * e;
    }
  }
  return mu * ee + 0.5 * lm * tr * tr;
}
static void P_le(int d, const double* G, double mu, double lm, double* P) {
  double tr = 0;
  for (int i = 0; i < d; ++i) tr += G[i * d + i];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) P[i * d + j] = mu * (G[i * d + j] + G[j * d + i]) + (i == j ? lm * tr : 0.0);
}

static void nh_prepare(const double* G, double* F, double* Fi, double* lnJ) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F[i * 3 + j] = G[i * 3 + j] + (i == j);
  *lnJ = log(inv_small(3, F, Fi));
}
static double psi_nh(const double* G, double mu, double lm) {
  double F[9], Fi[9], lnJ, I1 = 0;
  nh_prepare(G, F, Fi, &lnJ);
  for (int k = 0; k < 9; ++k) I1 += F[k] * F[k];
  return 0.5 * mu * (I1 - 3 - 2 * lnJ) + 0.5 * lm * lnJ * lnJ;
}
static void P_nh(const double* G, double mu, double lm, double* P) {
  double F[9], Fi[9], lnJ;
  nh_prepare(G, F, Fi, &lnJ);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) P[i * 3 + j] = mu * (F[i * 3 + j] - Fi[j * 3 + i]) + lm * lnJ * Fi[j * 3 + i];
}
static void dP_nh(const double* G, const double* dG, double mu, double lm, double* dP) {
  double F[9], Fi[9], lnJ, B[9], tr = 0;
  nh_prepare(G, F, Fi, &lnJ);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0;
      for (int k = 0; k < 3; ++k) t += Fi[i * 3 + k] * dG[k * 3 + j];
      B[i * 3 + j] = t;
      if (i == j) tr += t;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0;
      for (int k = 0; k < 3; ++k) t += B[j * 3 + k] * Fi[k * 3 + i];
      dP[i * 3 + j] = mu * dG[i * 3 + j] + (mu - lm * lnJ) * t + lm * tr * Fi[j * 3 + i];
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* mode 0: *out = E(u);  mode 1: out (N,dim) = dE/du;  mode 2: out = H(u) v.   returns 0 on success */
int oracle_fused(int kind, int material, int mode, double mu, double lm, int64_t n_nodes, int64_t n_elems,
                 const double* coords, const int32_t* conn, const double* u, const double* v, double* out) {
  int dim, npe, nq;
  elem_info(kind, &dim, &npe, &nq);
  if (material == NEO_HOOKEAN && dim != 3) return -1;
  const double wq = quad_weight(kind);
  double energy = 0.0;
  if (mode != MODE_ENERGY) memset(out, 0, sizeof(double) * n_nodes * dim);
#pragma omp parallel for schedule(static) reduction(+ : energy)
  for (int64_t e = 0; e < n_elems; ++e) {
    double X[24], U[24], V[24], Y[24] = {0}, dNdX[24], G[9], dG[9], P[9];
    const int32_t* nd = conn + e * npe;
    for (int n = 0; n < npe; ++n)
      for (int c = 0; c < dim; ++c) {
        X[n * dim + c] = coords[(int64_t)nd[n] * dim + c];
        U[n * dim + c] = u[(int64_t)nd[n] * dim + c];
        if (mode == MODE_HVP) V[n * dim + c] = v[(int64_t)nd[n] * dim + c];
      }
    for (int q = 0; q < nq; ++q) {
      const double W = geometry(kind, dim, npe, q, X, dNdX) * wq;
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
          double s = 0, t = 0;
          for (int n = 0; n < npe; ++n) {
            s += dNdX[j * npe + n] * U[n * dim + i];
            if (mode == MODE_HVP) t += dNdX[j * npe + n] * V[n * dim + i];
          }
          G[i * dim + j] = s;
          dG[i * dim + j] = t;
        }
      if (mode == MODE_ENERGY) {
        energy += W * (material == NEO_HOOKEAN ? psi_nh(G, mu, lm) : psi_le(dim, G, mu, lm));
        continue;
      }
      if (material == NEO_HOOKEAN) {
        if (mode == MODE_RESIDUAL) P_nh(G, mu, lm, P); else dP_nh(G, dG, mu, lm, P);
      } else {
        P_le(dim, mode == MODE_RESIDUAL ? G : dG, mu, lm, P);
      }
      for (int n = 0; n < npe; ++n)
        for (int i = 0; i < dim; ++i) {
          double s = 0;
          for (int j = 0; j < dim; ++j) s += P[i * dim + j] * dNdX[j * npe + n];
          Y[n * dim + i] += W * s;
        }
    }
    if (mode != MODE_ENERGY)
      for (int n = 0; n < npe; ++n)
        for (int i = 0; i < dim; ++i) {
#pragma omp atomic
          out[(int64_t)nd[n] * dim + i] += Y[n * dim + i];
        }
  }
  if (mode == MODE_ENERGY) *out = energy;
  return 0;
}

/* ---- Operator building blocks at the config sizes (test infrastructure for the full-size GPU parity tests) ------------
 * what 0: out (E, Q, nv, dim) = Operator.grad(u)            (tatva/operator.py:379-397 -> element/base.py:99-115)
 * what 1: out (N, nv)        += adjoint of grad applied to in (E, Q, nv, dim)   (the transposed gather / gradient)
 * what 2: out (E, Q)          = Operator.get_integration_weights()              (tatva/operator.py:172-192)
 * what 3: out (E, npe, nv)    = in[mesh.elements]                                (tatva/operator.py:221)
 * `in` is the nodal field (N, nv) for what 0 / 3 and the quadrature-shaped array for what 1.  returns 0 on success.      */
int oracle_blocks(int kind, int what, int nv, int64_t n_nodes, int64_t n_elems, const double* coords, const int32_t* conn,
                  const double* in, double* out) {
  int dim, npe, nq;
  elem_info(kind, &dim, &npe, &nq);
  if (what < 0 || what > 3 || nv < 1 || nv > 8) return -1;
  const double wq = quad_weight(kind);
  if (what == 1) memset(out, 0, sizeof(double) * n_nodes * nv);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_elems; ++e) {
    double X[24], dNdX[24];
    const int32_t* nd = conn + e * npe;
    if (what == 3) {
      for (int n = 0; n < npe; ++n)
        for (int c = 0; c < nv; ++c) out[(e * npe + n) * nv + c] = in[(int64_t)nd[n] * nv + c];
      continue;
    }
    for (int n = 0; n < npe; ++n)
      for (int c = 0; c < dim; ++c) X[n * dim + c] = coords[(int64_t)nd[n] * dim + c];
    for (int q = 0; q < nq; ++q) {
      const double det = geometry(kind, dim, npe, q, X, dNdX);
      if (what == 2) {
        out[e * nq + q] = det * wq;
      } else if (what == 0) {
        for (int c = 0; c < nv; ++c)
          for (int j = 0; j < dim; ++j) {
            double t = 0;
            for (int n = 0; n < npe; ++n) t += dNdX[j * npe + n] * in[(int64_t)nd[n] * nv + c];
            out[((e * nq + q) * nv + c) * dim + j] = t;
          }
      } else {
        for (int c = 0; c < nv; ++c)
          for (int n = 0; n < npe; ++n) {
            double t = 0;
            for (int j = 0; j < dim; ++j) t += in[((e * nq + q) * nv + c) * dim + j] * dNdX[j * npe + n];
#pragma omp atomic
            out[(int64_t)nd[n] * nv + c] += t;
          }
      }
    }
  }
  return 0;
}

/* ---- Compound (u, phi) two-field law of config 5 (builder-defined AT2 density; no counterpart in the reference, see
 * oracle/tatva_oracle.py NeoHookeanPhaseField).  Nodal state s (N,4) = [ux,uy,uz,phi] (compound/__init__.py:334-389);
 * the fields enter through Operator.grad (u, phi) and Operator.eval (phi): operator.py:358-397, element/base.py:95-115. */
static void shape_fn(int kind, int q, double* N) {
  if (kind == TET4) {
    for (int n = 0; n < 4; ++n) N[n] = 0.25; /* centroid: 1-xi-eta-zeta = xi = eta = zeta = 1/4  (base.py:448-466) */
  } else {
    const double a = 1.0 / sqrt(3.0);
    for (int n = 0; n < 8; ++n)
      N[n] = 0.125 * (1 + HEX_S[n][0] * a * HEX_S[q][0]) * (1 + HEX_S[n][1] * a * HEX_S[q][1]) * (1 + HEX_S[n][2] * a * HEX_S[q][2]);
  }
}

/* prm = {mu, lambda, Gc, ell, k}.  mode as oracle_fused; s, t, out are (N,4). */
int oracle_fused_pf(int kind, int mode, const double* prm, int64_t n_nodes, int64_t n_elems, const double* coords,
                    const int32_t* conn, const double* s, const double* t, double* out) {
  int dim, npe, nq;
  elem_info(kind, &dim, &npe, &nq);
  if (dim != 3) return -1;
  const double mu = prm[0], lm = prm[1], Gc = prm[2], ell = prm[3], kk = prm[4];
  const double wq = quad_weight(kind);
  double energy = 0.0;
  if (mode != MODE_ENERGY) memset(out, 0, sizeof(double) * n_nodes * 4);
#pragma omp parallel for schedule(static) reduction(+ : energy)
  for (int64_t e = 0; e < n_elems; ++e) {
    double X[24], S[32], T[32], Y[32] = {0}, dNdX[24], N[8];
    const int32_t* nd = conn + e * npe;
    for (int n = 0; n < npe; ++n) {
      for (int c = 0; c < 3; ++c) X[n * 3 + c] = coords[(int64_t)nd[n] * 3 + c];
      for (int c = 0; c < 4; ++c) {
        S[n * 4 + c] = s[(int64_t)nd[n] * 4 + c];
        if (mode == MODE_HVP) T[n * 4 + c] = t[(int64_t)nd[n] * 4 + c];
      }
    }
    for (int q = 0; q < nq; ++q) {
      const double W = geometry(kind, dim, npe, q, X, dNdX) * wq;
      shape_fn(kind, q, N);
      double G[9], dG[9], gphi[3], dgphi[3], phi = 0, dphi = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double a = 0, b = 0;
          for (int n = 0; n < npe; ++n) {
            a += dNdX[j * npe + n] * S[n * 4 + i];
            if (mode == MODE_HVP) b += dNdX[j * npe + n] * T[n * 4 + i];
          }
          G[i * 3 + j] = a;
          dG[i * 3 + j] = b;
        }
      for (int j = 0; j < 3; ++j) {
        double a = 0, b = 0;
        for (int n = 0; n < npe; ++n) {
          a += dNdX[j * npe + n] * S[n * 4 + 3];
          if (mode == MODE_HVP) b += dNdX[j * npe + n] * T[n * 4 + 3];
        }
        gphi[j] = a;
        dgphi[j] = b;
      }
      for (int n = 0; n < npe; ++n) {
        phi += N[n] * S[n * 4 + 3];
        if (mode == MODE_HVP) dphi += N[n] * T[n * 4 + 3];
      }
      const double g = (1 - phi) * (1 - phi) + kk, dg = -2 * (1 - phi);
      const double psi = psi_nh(G, mu, lm);
      if (mode == MODE_ENERGY) {
        energy += W * (g * psi + Gc * (phi * phi / (2 * ell) + 0.5 * ell * (gphi[0] * gphi[0] + gphi[1] * gphi[1] + gphi[2] * gphi[2])));
        continue;
      }
      double P[9], A[9], b, c[3];
      P_nh(G, mu, lm, P);
      if (mode == MODE_RESIDUAL) {
        for (int k = 0; k < 9; ++k) A[k] = g * P[k];
        b = dg * psi + Gc * phi / ell;
        for (int j = 0; j < 3; ++j) c[j] = Gc * ell * gphi[j];
      } else {
        double dP[9], PdG = 0;
        dP_nh(G, dG, mu, lm, dP);
        for (int k = 0; k < 9; ++k) PdG += P[k] * dG[k];
        for (int k = 0; k < 9; ++k) A[k] = g * dP[k] + dg * dphi * P[k];
        b = dg * PdG + 2 * dphi * psi + Gc * dphi / ell;
        for (int j = 0; j < 3; ++j) c[j] = Gc * ell * dgphi[j];
      }
      for (int n = 0; n < npe; ++n) {
        for (int i = 0; i < 3; ++i) {
          double a = 0;
          for (int j = 0; j < 3; ++j) a += A[i * 3 + j] * dNdX[j * npe + n];
          Y[n * 4 + i] += W * a;
        }
        double a = b * N[n];
        for (int j = 0; j < 3; ++j) a += c[j] * dNdX[j * npe + n];
        Y[n * 4 + 3] += W * a;
      }
    }
    if (mode != MODE_ENERGY)
      for (int n = 0; n < npe; ++n)
        for (int i = 0; i < 4; ++i) {
#pragma omp atomic
          out[(int64_t)nd[n] * 4 + i] += Y[n * 4 + i];
        }
  }
  if (mode == MODE_ENERGY) *out = energy;
  return 0;
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
