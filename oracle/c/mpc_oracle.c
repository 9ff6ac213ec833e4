/*
 * mpc_oracle.c -- plain-C CPU restatement of the convex-MPC stance solve.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: restates ``mpc_osqp.ConvexMpc::ComputeContactForces`` of motion_imitation==0.0.5
 * (third-party C++/Eigen/OSQP, absent from /root/reference; call site
 * robot_gym/controllers/mpc/mpc_controller.py:47-56,105) exactly as oracle/convex_mpc.py does, in the
 * reference's own dense formulation: A/B matrices, Pade scaling-and-squaring matrix exponential of
 * [[A,B],[0,0]]*dt (what Eigen's MatrixBase::exp() does), stacked A_qp / B_qp, dense
 * P = 2 (B_qp^T L B_qp + alpha I), q = 2 B_qp^T L (A_qp x0 - x_ref), 5 pyramid rows per foot per step.
 * The QP is solved by a dense Mehrotra predictor-corrector interior point followed by an exact
 * active-set polish (the reference uses OSQP ADMM + polish; OSQP is not available offline -- both
 * converge to the unique optimum of the strictly convex QP).  Swing feet (0 <= C f <= 0) are eliminated before the solve.
 *
 * Roles: (1) fast checker for the 4096-env GPU parity test, pinned against oracle/convex_mpc.py;
 * (2) the CPU baseline bench.py times ("port", one env per OpenMP thread iteration).
 * It shares no code and no algebraic shortcut with the CUDA kernels.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define S13 13
#define MAXH 20
#define MAXK 4

typedef struct rgo_params {
  double mass;
  double inertia[9];
  int num_legs;
  int horizon;
  double dt;
  double weights[13];
  double alpha;
  double mu[4];
  double gravity;
  double fz_max;
  double fz_min;
} rgo_params;

/* ------------------------------------------------------------------ small dense helpers */
static void matmul(const double* a, const double* b, double* c, int m, int k, int n) {
  for (int i = 0; i < m; ++i) {
    double* ci = c + (size_t)i * n;
    for (int j = 0; j < n; ++j) ci[j] = 0.0;
    for (int l = 0; l < k; ++l) {
      const double ail = a[(size_t)i * k + l];
      if (ail == 0.0) continue;
      const double* bl = b + (size_t)l * n;
      for (int j = 0; j < n; ++j) ci[j] += ail * bl[j];
    }
  }
}

/* solve A X = B in place (LU with partial pivoting); A n x n, B n x m; returns 0 on success */
static int lu_solve(double* a, double* b, int n, int m) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
    if (best == 0.0) return -1;
    if (piv != c) {
      for (int j = 0; j < n; ++j) { double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t; }
      for (int j = 0; j < m; ++j) { double t = b[c * m + j]; b[c * m + j] = b[piv * m + j]; b[piv * m + j] = t; }
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = a[r * n + c] / a[c * n + c];
      if (f == 0.0) continue;
      for (int j = c; j < n; ++j) a[r * n + j] -= f * a[c * n + j];
      for (int j = 0; j < m; ++j) b[r * m + j] -= f * b[c * m + j];
    }
  }
  for (int c = n - 1; c >= 0; --c)
    for (int j = 0; j < m; ++j) {
      double v = b[c * m + j];
      for (int l = c + 1; l < n; ++l) v -= a[c * n + l] * b[l * m + j];
      b[c * m + j] = v / a[c * n + c];
    }
  return 0;
}

/* Matrix exponential by Pade approximation with scaling and squaring (Higham 2005), the algorithm
 * behind Eigen's MatrixBase::exp() used in CalculateExponentials. */
static void expm_pade(const double* a_in, double* out, int n) {
  static const double theta[5] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1,
                                  2.097847961257068e0, 5.371920351148152e0};
  static const double b3[] = {120., 60., 12., 1.};
  static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
  static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
  static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
  static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                               129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                               40840800., 960960., 16380., 182., 1.};
  const size_t nn = (size_t)n * n;
  double* a = (double*)malloc(9 * nn * sizeof(double));
  double *a2 = a + nn, *a4 = a2 + nn, *a6 = a4 + nn, *u = a6 + nn, *v = u + nn, *t1 = v + nn, *t2 = t1 + nn, *a8 = t2 + nn;
  memcpy(a, a_in, nn * sizeof(double));
  double norm1 = 0.0;
  for (int j = 0; j < n; ++j) { double s = 0.0; for (int i = 0; i < n; ++i) s += fabs(a[i * n + j]); if (s > norm1) norm1 = s; }
  int squarings = 0;
  int degree = 13;
  if (norm1 <= theta[0]) degree = 3; else if (norm1 <= theta[1]) degree = 5; else if (norm1 <= theta[2]) degree = 7;
  else if (norm1 <= theta[3]) degree = 9;
  else {
    squarings = (int)fmax(0.0, ceil(log2(norm1 / theta[4])));
    const double sc = ldexp(1.0, -squarings);
    for (size_t i = 0; i < nn; ++i) a[i] *= sc;
  }
  matmul(a, a, a2, n, n, n);
  matmul(a2, a2, a4, n, n, n);
  matmul(a4, a2, a6, n, n, n);
  /* U = A * (odd part), V = even part */
  if (degree <= 9) {
    const double* b = degree == 3 ? b3 : degree == 5 ? b5 : degree == 7 ? b7 : b9;
    if (degree == 9) matmul(a6, a2, a8, n, n, n);
    const double* pw[5] = {NULL, a2, a4, a6, a8};
    for (size_t i = 0; i < nn; ++i) { t1[i] = 0.0; v[i] = 0.0; }
    for (int i = 0; i < n; ++i) { t1[i * n + i] = b[1]; v[i * n + i] = b[0]; }
    for (int k = 1; 2 * k <= degree; ++k)
      for (size_t i = 0; i < nn; ++i) { t1[i] += b[2 * k + 1] * pw[k][i]; v[i] += b[2 * k] * pw[k][i]; }
    matmul(a, t1, u, n, n, n);
  } else {
    for (size_t i = 0; i < nn; ++i) t1[i] = b13[13] * a6[i] + b13[11] * a4[i] + b13[9] * a2[i];
    matmul(a6, t1, t2, n, n, n);
    for (size_t i = 0; i < nn; ++i) t2[i] += b13[7] * a6[i] + b13[5] * a4[i] + b13[3] * a2[i];
    for (int i = 0; i < n; ++i) t2[i * n + i] += b13[1];
    matmul(a, t2, u, n, n, n);
    for (size_t i = 0; i < nn; ++i) t1[i] = b13[12] * a6[i] + b13[10] * a4[i] + b13[8] * a2[i];
    matmul(a6, t1, v, n, n, n);
    for (size_t i = 0; i < nn; ++i) v[i] += b13[6] * a6[i] + b13[4] * a4[i] + b13[2] * a2[i];
    for (int i = 0; i < n; ++i) v[i * n + i] += b13[0];
  }
  /* (V - U) X = (V + U) */
  for (size_t i = 0; i < nn; ++i) { t1[i] = v[i] - u[i]; t2[i] = v[i] + u[i]; }
  lu_solve(t1, t2, n, n);
  for (int s = 0; s < squarings; ++s) { matmul(t2, t2, t1, n, n, n); memcpy(t2, t1, nn * sizeof(double)); }
  memcpy(out, t2, nn * sizeof(double));
  free(a);
}

static void rot_x(double a, double* r) { double c = cos(a), s = sin(a); double m[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(r, m, sizeof(m)); }
static void rot_y(double a, double* r) { double c = cos(a), s = sin(a); double m[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(r, m, sizeof(m)); }
static void rot_z(double a, double* r) { double c = cos(a), s = sin(a); double m[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; memcpy(r, m, sizeof(m)); }

static int inv3(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (det == 0.0) return -1;
  const double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return 0;
}

/* in-place Cholesky (lower), full row-major storage; returns -1 if not positive definite */
static int cholesky(double* a, int n) {
  for (int j = 0; j < n; ++j) {
    double d = a[j * n + j];
    for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    a[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = a[i * n + j];
      const double *ri = a + (size_t)i * n, *rj = a + (size_t)j * n;
      for (int k = 0; k < j; ++k) v -= ri[k] * rj[k];
      a[i * n + j] = v / d;
    }
  }
  return 0;
}

static void chol_solve(const double* l, double* b, int n) {
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= l[i * n + k] * b[k];
    b[i] = v / l[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= l[k * n + i] * b[k];
    b[i] = v / l[i * n + i];
  }
}

/* ------------------------------------------------------------------ one env */
size_t rgo_scratch_doubles(int horizon, int num_legs);
typedef struct work {
  double *b_qp, *p_full, *q_full, *p, *q, *phi, *g, *hv, *x, *s, *lam, *rd, *rhs, *dx, *dxa, *ds, *dl, *dsa, *dla, *tmp, *xbest;
} work;

static size_t work_doubles(int h, int k) {
  const size_t n = 3 * (size_t)k * h, m = 10 * (size_t)k * h, ns = (size_t)S13 * h;
  return ns * n + 3 * n * n + 4 * n + m * 3 + 7 * m + 8 * n + 64;
}

/* Solves one QP; out = -(solution) (3*k*h doubles).  Returns the interior-point iteration count. */
int rgo_compute_contact_forces(const rgo_params* p, const double* com_vel, const double* rpy, const double* ang_vel,
                               const int* contacts, const double* feet_base, const double* des_pos,
                               const double* des_vel, const double* des_rpy, const double* des_w,
                               const double* com_pos /* 3 or NULL */, double* out, double* scratch) {
  const int h = p->horizon, k = p->num_legs, m3 = 3 * k, n_all = m3 * h, ns = S13 * h;
  const double dt = p->dt, g = p->gravity;
  memset(out, 0, sizeof(double) * n_all);
  int n_stance = 0;
  for (int i = 0; i < k; ++i) n_stance += contacts[i] != 0;
  if (n_stance == 0) return 0;

  /* foot positions in the world frame: R = Rx Ry Rz (sic) */
  double rx[9], ry[9], rz[9], t33[9], rfeet[9], rbody[9];
  rot_x(rpy[0], rx); rot_y(rpy[1], ry); rot_z(rpy[2], rz);
  matmul(rx, ry, t33, 3, 3, 3); matmul(t33, rz, rfeet, 3, 3, 3);
  matmul(rz, ry, t33, 3, 3, 3); matmul(t33, rx, rbody, 3, 3, 3);
  double feet_w[MAXK][3];
  double zsum = 0.0;
  for (int i = 0; i < k; ++i) {
    for (int r = 0; r < 3; ++r)
      feet_w[i][r] = rfeet[3 * r] * feet_base[3 * i] + rfeet[3 * r + 1] * feet_base[3 * i + 1] + rfeet[3 * r + 2] * feet_base[3 * i + 2];
    if (contacts[i]) zsum += feet_w[i][2];
  }
  const double com_z = com_pos ? com_pos[2] : fabs(zsum / n_stance);

  double x0[S13] = {rpy[0], rpy[1], rpy[2], 0.0, 0.0, com_z, ang_vel[0], ang_vel[1], ang_vel[2],
                    com_vel[0], com_vel[1], com_vel[2], -g};

  /* continuous A, B and the exponential of [[A,B],[0,0]] dt */
  const int nab = S13 + m3;
  double ab[(S13 + 3 * MAXK) * (S13 + 3 * MAXK)], abexp[(S13 + 3 * MAXK) * (S13 + 3 * MAXK)];
  memset(ab, 0, sizeof(ab));
  {
    const double cy = cos(rpy[2]), sy = sin(rpy[2]), cp = cos(rpy[1]), tp = tan(rpy[1]);
    const double tm[9] = {cy / cp, sy / cp, 0, -sy, cy, 0, cy * tp, sy * tp, 1};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) ab[r * nab + 6 + c] = tm[3 * r + c] * dt;
    ab[3 * nab + 9] = ab[4 * nab + 10] = ab[5 * nab + 11] = dt;
    ab[11 * nab + 12] = dt;
    double inv_i[9] = {0}, iw[9];
    inv3(p->inertia, inv_i);
    matmul(rbody, inv_i, t33, 3, 3, 3);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      double v = 0.0; for (int c = 0; c < 3; ++c) v += t33[3 * a + c] * rbody[3 * b + c];
      iw[3 * a + b] = v;
    }
    for (int i = 0; i < k; ++i) {
      const double* r = feet_w[i];
      const double sk[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
      double blk[9];
      matmul(iw, sk, blk, 3, 3, 3);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) ab[(6 + a) * nab + S13 + 3 * i + b] = blk[3 * a + b] * dt;
      for (int a = 0; a < 3; ++a) ab[(9 + a) * nab + S13 + 3 * i + a] = dt / p->mass;
    }
  }
  expm_pade(ab, abexp, nab);
  double a_exp[S13 * S13], b_exp[S13 * 3 * MAXK];
  for (int r = 0; r < S13; ++r) {
    for (int c = 0; c < S13; ++c) a_exp[r * S13 + c] = abexp[r * nab + c];
    for (int c = 0; c < m3; ++c) b_exp[r * m3 + c] = abexp[r * nab + S13 + c];
  }

  /* carve the scratch buffer */
  double* w = scratch;
  double* b_qp = w; w += (size_t)ns * n_all;
  double* p_full = w; w += (size_t)n_all * n_all;
  double* q_full = w; w += n_all;
  /* A_qp x0 (free response) and A^i B blocks */
  double anb[MAXH][S13 * 3 * MAXK];
  double free_resp[MAXH * S13];
  {
    double apow[S13 * S13], nxt[S13 * S13], xs[S13], xn[S13];
    memcpy(anb[0], b_exp, sizeof(double) * S13 * m3);
    memcpy(apow, a_exp, sizeof(apow));
    memcpy(xs, x0, sizeof(xs));
    for (int i = 0; i < h; ++i) {
      matmul(a_exp, xs, xn, S13, S13, 1);
      memcpy(xs, xn, sizeof(xs));
      memcpy(free_resp + i * S13, xs, sizeof(xs));
      if (i + 1 < h) { matmul(a_exp, anb[i], anb[i + 1], S13, S13, m3); }
    }
    (void)apow; (void)nxt;
  }
  memset(b_qp, 0, sizeof(double) * (size_t)ns * n_all);
  for (int i = 0; i < h; ++i)
    for (int j = 0; j <= i; ++j)
      for (int r = 0; r < S13; ++r)
        memcpy(b_qp + (size_t)(i * S13 + r) * n_all + j * m3, anb[i - j] + r * m3, sizeof(double) * m3);
  /* state error and q = 2 B^T L (A x0 - x_ref);  P = 2 (B^T L B + alpha I) */
  double sd[MAXH * S13];
  for (int i = 0; i < h; ++i) {
    double ref[S13] = {des_rpy[0], des_rpy[1], rpy[2] + dt * (i + 1) * des_w[2], dt * (i + 1) * des_vel[0],
                       dt * (i + 1) * des_vel[1], des_pos[2], 0.0, 0.0, des_w[2], des_vel[0], des_vel[1], 0.0, -g};
    for (int r = 0; r < S13; ++r) sd[i * S13 + r] = p->weights[r] * (free_resp[i * S13 + r] - ref[r]);
  }
  for (int c = 0; c < n_all; ++c) {
    double v = 0.0;
    for (int r = 0; r < ns; ++r) v += b_qp[(size_t)r * n_all + c] * sd[r];
    q_full[c] = 2.0 * v;
  }
  for (int a = 0; a < n_all; ++a)
    for (int b = 0; b <= a; ++b) {
      double v = 0.0;
      const int first = (a / m3) * S13;       /* B_qp is block lower triangular */
      for (int r = first; r < ns; ++r) v += p->weights[r % S13] * b_qp[(size_t)r * n_all + a] * b_qp[(size_t)r * n_all + b];
      v = 2.0 * (v + (a == b ? p->alpha : 0.0));
      p_full[(size_t)a * n_all + b] = p_full[(size_t)b * n_all + a] = v;
    }

  /* eliminate swing feet */
  int idx[3 * MAXK * MAXH];
  int n = 0;
  for (int t = 0; t < h; ++t)
    for (int l = 0; l < k; ++l)
      if (contacts[l]) for (int d = 0; d < 3; ++d) idx[n++] = (t * k + l) * 3 + d;
  const int nblk = n / 3, m = 10 * nblk;
  double* pm = w; w += (size_t)n * n;
  double* phi = w; w += (size_t)n * n;
  double* qv = w; w += n;
  double* x = w; w += n; double* xbest = w; w += n;
  double* rd = w; w += n; double* rhs = w; w += n; double* dxa = w; w += n; double* dx = w; w += n; double* pux = w; w += n;
  double* s = w; w += m; double* lam = w; w += m; double* dsa = w; w += m; double* dla = w; w += m;
  double* ds = w; w += m; double* dl = w; w += m; double* wv = w; w += m;
  for (int a = 0; a < n; ++a) {
    qv[a] = q_full[idx[a]];
    for (int b = 0; b < n; ++b) pm[(size_t)a * n + b] = p_full[(size_t)idx[a] * n_all + idx[b]];
  }
  const double* mu = p->mu;
  const double grow[5][3] = {{-1, 0, mu[0]}, {1, 0, mu[1]}, {0, -1, mu[2]}, {0, 1, mu[3]}, {0, 0, 1}};
  const double big_u = (mu[0] + 1.0) * p->fz_max;
  const double hup[5] = {big_u, big_u, big_u, big_u, p->fz_max};
  const double lo[5] = {0, 0, 0, 0, p->fz_min};

  double qscale = 1.0;
  for (int a = 0; a < n; ++a) if (fabs(qv[a]) > qscale) qscale = fabs(qv[a]);
  const double fz0 = sqrt(fmax(p->fz_min, 1e-3 * p->fz_max) * p->fz_max);
  for (int b = 0; b < nblk; ++b) {
    x[3 * b] = x[3 * b + 1] = 0.0; x[3 * b + 2] = fz0;
    for (int r = 0; r < 5; ++r) {
      const double c = grow[r][0] * x[3 * b] + grow[r][1] * x[3 * b + 1] + grow[r][2] * x[3 * b + 2];
      s[10 * b + r] = hup[r] - c;
      s[10 * b + 5 + r] = c - lo[r];
    }
    for (int r = 0; r < 10; ++r) lam[10 * b + r] = 0.1 * qscale / s[10 * b + r];
  }
  memcpy(xbest, x, sizeof(double) * n);
  double best = 1e300, prev = 1e300;
  int stall = 0, it = 0;
  const double tol = 1e-12;
  for (;;) {
    matmul(pm, x, pux, n, n, 1);
    double rdmax = 0.0, sl = 0.0;
    for (int b = 0; b < nblk; ++b) {
      double e[5];
      for (int r = 0; r < 5; ++r) e[r] = lam[10 * b + r] - lam[10 * b + 5 + r];
      const double gl[3] = {e[1] - e[0], e[3] - e[2], mu[0] * e[0] + mu[1] * e[1] + mu[2] * e[2] + mu[3] * e[3] + e[4]};
      for (int d = 0; d < 3; ++d) { rd[3 * b + d] = pux[3 * b + d] + qv[3 * b + d] + gl[d]; if (fabs(rd[3 * b + d]) > rdmax) rdmax = fabs(rd[3 * b + d]); }
      for (int r = 0; r < 10; ++r) sl += s[10 * b + r] * lam[10 * b + r];
    }
    const double mu_c = sl / m;
    const double res = fmax(rdmax, mu_c) / qscale;
    if (res < best) { best = res; memcpy(xbest, x, sizeof(double) * n); }
    if (res < tol) break;
    stall = (res > 0.9 * prev && res < 1e-7) ? stall + 1 : 0;
    prev = res;
    if (it >= 60 || stall >= 4) break;
    ++it;
    /* Phi = P + G^T D G */
    memcpy(phi, pm, sizeof(double) * (size_t)n * n);
    for (int b = 0; b < nblk; ++b) {
      double dd[5];
      for (int r = 0; r < 5; ++r) dd[r] = lam[10 * b + r] / s[10 * b + r] + lam[10 * b + 5 + r] / s[10 * b + 5 + r];
      for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
        double v = 0.0; for (int r = 0; r < 5; ++r) v += dd[r] * grow[r][a] * grow[r][c];
        phi[(size_t)(3 * b + a) * n + 3 * b + c] += v;
      }
    }
    if (cholesky(phi, n) != 0) break;
    for (int a = 0; a < n; ++a) dxa[a] = -(pux[a] + qv[a]);
    chol_solve(phi, dxa, n);
    double amax = 1.0;
    for (int b = 0; b < nblk; ++b) {
      double c5[5];
      for (int r = 0; r < 5; ++r) c5[r] = grow[r][0] * dxa[3 * b] + grow[r][1] * dxa[3 * b + 1] + grow[r][2] * dxa[3 * b + 2];
      for (int r = 0; r < 10; ++r) {
        const int j = 10 * b + r;
        dsa[j] = r < 5 ? -c5[r] : c5[r - 5];
        dla[j] = -lam[j] - lam[j] * dsa[j] / s[j];
        if (dsa[j] < 0 && -s[j] / dsa[j] < amax) amax = -s[j] / dsa[j];
        if (dla[j] < 0 && -lam[j] / dla[j] < amax) amax = -lam[j] / dla[j];
      }
    }
    double mu_aff = 0.0;
    for (int j = 0; j < m; ++j) mu_aff += (s[j] + amax * dsa[j]) * (lam[j] + amax * dla[j]);
    mu_aff /= m;
    const double sig = pow(mu_aff / mu_c, 3.0), sigmu = sig * mu_c;
    for (int j = 0; j < m; ++j) wv[j] = (s[j] * lam[j] + dsa[j] * dla[j] - sigmu) / s[j];
    for (int b = 0; b < nblk; ++b) {
      double e[5];
      for (int r = 0; r < 5; ++r) e[r] = wv[10 * b + r] - wv[10 * b + 5 + r];
      const double gw[3] = {e[1] - e[0], e[3] - e[2], mu[0] * e[0] + mu[1] * e[1] + mu[2] * e[2] + mu[3] * e[3] + e[4]};
      for (int d = 0; d < 3; ++d) dx[3 * b + d] = -rd[3 * b + d] + gw[d];
    }
    chol_solve(phi, dx, n);
    double step = 1e30;
    for (int b = 0; b < nblk; ++b) {
      double c5[5];
      for (int r = 0; r < 5; ++r) c5[r] = grow[r][0] * dx[3 * b] + grow[r][1] * dx[3 * b + 1] + grow[r][2] * dx[3 * b + 2];
      for (int r = 0; r < 10; ++r) {
        const int j = 10 * b + r;
        ds[j] = r < 5 ? -c5[r] : c5[r - 5];
        dl[j] = (-(s[j] * lam[j] + dsa[j] * dla[j] - sigmu) - lam[j] * ds[j]) / s[j];
        if (ds[j] < 0 && -s[j] / ds[j] < step) step = -s[j] / ds[j];
        if (dl[j] < 0 && -lam[j] / dl[j] < step) step = -lam[j] / dl[j];
      }
    }
    step = fmin(1.0, 0.99 * step);
    for (int a = 0; a < n; ++a) x[a] += step * dx[a];
    for (int j = 0; j < m; ++j) { s[j] += step * ds[j]; lam[j] += step * dl[j]; }
  }
  /* Active-set polish (the counterpart of OSQP's polish and of oracle/convex_mpc.py's refinement): solve the
   * equality-constrained QP on the rows the interior point marks active exactly,
   *     x = x_unc - Y y,   Y = P^-1 C_a^T,   (C_a Y) y = C_a x_unc - b_a,
   * add violated rows, drop rows whose multiplier has the wrong sign, repeat.  Accepted only when a round
   * changes nothing (a KKT point of a strictly convex QP = the optimum); otherwise the interior-point
   * iterate stands.  phi / dxa / dx / rhs / wv are free here and reused as scratch. */
  {
    int* side = (int*)malloc(sizeof(int) * (size_t)(5 * nblk));          /* +1 upper, -1 lower, 0 free */
    int* rows = (int*)malloc(sizeof(int) * (size_t)(5 * nblk));
    double* lfac = phi;                                                    /* Cholesky factor of P */
    double* xunc = dxa;
    double* xp = dx;
    memcpy(lfac, pm, sizeof(double) * (size_t)n * n);
    int ok = cholesky(lfac, n) == 0;
    if (ok) {
      for (int a = 0; a < n; ++a) xunc[a] = -qv[a];
      chol_solve(lfac, xunc, n);
      for (int b = 0; b < nblk; ++b)
        for (int r = 0; r < 5; ++r) {
          const double c = grow[r][0] * xbest[3 * b] + grow[r][1] * xbest[3 * b + 1] + grow[r][2] * xbest[3 * b + 2];
          const double span_hi = fmax(1.0, fabs(hup[r])), span_lo = fmax(1.0, fabs(lo[r]));
          side[5 * b + r] = 0;
          if (lam[10 * b + r] > s[10 * b + r] && hup[r] - c < 1e-5 * span_hi) side[5 * b + r] = 1;
          if (lam[10 * b + 5 + r] > s[10 * b + 5 + r] && c - lo[r] < 1e-5 * fmax(span_hi, span_lo)) side[5 * b + r] = -1;
        }
    }
    const double feas_tol = 1e-11 * fmax(1.0, big_u);
    for (int round = 0; ok && round < 20; ++round) {
      int ma = 0;
      for (int j = 0; j < 5 * nblk; ++j) if (side[j]) rows[ma++] = j;
      double* ymat = (double*)malloc(sizeof(double) * (size_t)(ma > 0 ? ma : 1) * n);      /* rows of Y^T */
      double* smat = (double*)malloc(sizeof(double) * (size_t)(ma > 0 ? ma : 1) * (ma > 0 ? ma : 1));
      double* yv = (double*)malloc(sizeof(double) * (size_t)(ma > 0 ? ma : 1));
      for (int i = 0; i < ma; ++i) {
        const int b = rows[i] / 5, r = rows[i] % 5;
        double* col = ymat + (size_t)i * n;
        memset(col, 0, sizeof(double) * n);
        for (int d = 0; d < 3; ++d) col[3 * b + d] = grow[r][d];
        chol_solve(lfac, col, n);
      }
      for (int i = 0; i < ma; ++i) {
        const int b = rows[i] / 5, r = rows[i] % 5;
        for (int j = 0; j < ma; ++j) {
          const double* col = ymat + (size_t)j * n;
          smat[(size_t)i * ma + j] = grow[r][0] * col[3 * b] + grow[r][1] * col[3 * b + 1] + grow[r][2] * col[3 * b + 2];
        }
        const double target = side[rows[i]] > 0 ? hup[r] : lo[r];
        yv[i] = grow[r][0] * xunc[3 * b] + grow[r][1] * xunc[3 * b + 1] + grow[r][2] * xunc[3 * b + 2] - target;
      }
      if (ma > 0) {
        if (cholesky(smat, ma) != 0) { ok = 0; free(ymat); free(smat); free(yv); break; }   /* dependent rows: keep the IPM point */
        chol_solve(smat, yv, ma);
      }
      memcpy(xp, xunc, sizeof(double) * n);
      for (int i = 0; i < ma; ++i) {
        const double* col = ymat + (size_t)i * n;
        for (int a = 0; a < n; ++a) xp[a] -= yv[i] * col[a];
      }
      int changed = 0;
      double ymax = 1.0;
      for (int i = 0; i < ma; ++i) if (fabs(yv[i]) > ymax) ymax = fabs(yv[i]);
      for (int b = 0; b < nblk; ++b)
        for (int r = 0; r < 5; ++r) {
          if (side[5 * b + r]) continue;
          const double c = grow[r][0] * xp[3 * b] + grow[r][1] * xp[3 * b + 1] + grow[r][2] * xp[3 * b + 2];
          if (c - hup[r] > feas_tol) { side[5 * b + r] = 1; changed = 1; }
          else if (lo[r] - c > feas_tol) { side[5 * b + r] = -1; changed = 1; }
        }
      for (int i = 0; i < ma; ++i)
        if (side[rows[i]] * yv[i] < -1e-12 * ymax) { side[rows[i]] = 0; changed = 1; }
      free(ymat); free(smat); free(yv);
      if (!changed) { memcpy(xbest, xp, sizeof(double) * n); break; }
      if (round == 19) ok = 0;
    }
    free(side); free(rows);
  }
  for (int a = 0; a < n; ++a) out[idx[a]] = -xbest[a];
  (void)rhs;
  return it;
}


size_t rgo_scratch_doubles(int horizon, int num_legs) { return work_doubles(horizon, num_legs) + 4096; }

/* Batch driver over float32 env-major arrays (the layout the C ABI of the CUDA library takes).
 * forces_out: [n,12] first-step forces (float32).  Returns total interior-point iterations.
 * Threads: plain pthreads pulling chunks of 8 envs from a shared counter (libgomp is not in the image). */
typedef struct batch_job {
  const rgo_params* p;
  int n_env;
  const float *com_vel, *rpy, *rpy_rate, *feet, *command;
  const unsigned char* contacts;
  double desired_height;
  float *forces_out, *horizon_out;
  int next;                 /* next env chunk, updated atomically */
  long total_iters;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* job = (batch_job*)arg;
  const rgo_params* p = job->p;
  const int h = p->horizon, k = p->num_legs;
  double* scratch = (double*)malloc(sizeof(double) * rgo_scratch_doubles(h, k));
  double* out = (double*)malloc(sizeof(double) * 3 * k * h);
  long iters = 0;
  for (;;) {
    const int begin = __atomic_fetch_add(&job->next, 8, __ATOMIC_RELAXED);
    if (begin >= job->n_env) break;
    const int end = begin + 8 < job->n_env ? begin + 8 : job->n_env;
    for (int e = begin; e < end; ++e) {
      double cv[3], r3[3], w3[3], ft[12];
      int ct[4];
      for (int i = 0; i < 3; ++i) { cv[i] = job->com_vel[3 * e + i]; r3[i] = job->rpy[3 * e + i]; w3[i] = job->rpy_rate[3 * e + i]; }
      for (int i = 0; i < 12; ++i) ft[i] = job->feet[12 * e + i];
      for (int i = 0; i < 4; ++i) ct[i] = job->contacts[4 * e + i];
      const double dpos[3] = {0, 0, job->desired_height}, dvel[3] = {job->command[3 * e], job->command[3 * e + 1], 0.0};
      const double drpy[3] = {0, 0, 0}, dw[3] = {0, 0, job->command[3 * e + 2]};
      iters += rgo_compute_contact_forces(p, cv, r3, w3, ct, ft, dpos, dvel, drpy, dw, NULL, out, scratch);
      for (int i = 0; i < 12; ++i) job->forces_out[12 * e + i] = (float)out[i];
      if (job->horizon_out) for (int i = 0; i < 3 * k * h; ++i) job->horizon_out[(size_t)e * 3 * k * h + i] = (float)out[i];
    }
  }
  __atomic_fetch_add(&job->total_iters, iters, __ATOMIC_RELAXED);
  free(scratch);
  free(out);
  return NULL;
}

long rgo_batch(const rgo_params* p, int n_env, const float* com_vel, const float* rpy, const float* rpy_rate,
               const unsigned char* contacts, const float* feet, const float* command, double desired_height,
               float* forces_out, float* horizon_out /* [n,h,12] or NULL */, int n_threads) {
  batch_job job = {p, n_env, com_vel, rpy, rpy_rate, feet, command, contacts, desired_height, forces_out, horizon_out, 0, 0};
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t tid[256];
  for (int t = 1; t < n_threads; ++t) pthread_create(&tid[t], NULL, batch_worker, &job);
  batch_worker(&job);
  for (int t = 1; t < n_threads; ++t) pthread_join(tid[t], NULL);
  return job.total_iters;
}
