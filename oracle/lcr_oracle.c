/* lcr_oracle.c -- CPU float64 ORACLE for the env.step() hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (gym_lowcostrobot_b200 + liblcrsim.so) never does.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in the third-party `mujoco`
 * package (reference setup.py:11, `mujoco>=3.0`; the MJCF needs >= 3.1.3), which is neither vendored
 * under /root/reference nor installable here, and the reference's only test
 * (tests/test_env.py:8-13, gymnasium check_env) pins no numbers.  This file therefore restates
 *   (1) the reference's own Python glue line by line (file:line cited at each function), and
 *   (2) MuJoCo's published forward-dynamics pipeline (mj_step = mj_forward + implicitfast
 *       integration: kinematics, CRB + armature, collision, soft-constraint construction with
 *       solref/solimp impedance, elliptic-cone primal Newton solver, position actuators with
 *       per-joint force range) specialised to the model family of
 *       assets/low_cost_robot_6dof/follower.xml + one scene file.
 * (1) IS PINNED TO THE REFERENCE: tools/make_reference_glue_golden.py imports the unmodified reference env classes
 * (with stand-ins for the two absent third-party packages and n_substeps=0) and records reset sampling, action maps,
 * IK iterations, observations, rewards, success flags and PushCubeLoop's goal switching; tests/test_reference_glue.py
 * replays those fixtures through this file (to 1e-9) and through the CUDA path.
 * (2) is pinned by analytic known-answer tests (tests/test_oracle_known_answers.py) and by a second, structurally independent
 * numpy restatement built from the raw MJCF numbers by a separate reader (tests/test_independent_pipeline.py: the contact
 * SET by brute force; tests/test_independent_dynamics.py: one mj_step on all three scene classes), not by MuJoCo outputs.
 * tools/dump_mujoco_golden.py produces MuJoCo fixtures wherever MuJoCo exists; tests/test_golden_mujoco.py compares this
 * file and the CUDA path with them and reports which of the named switches (orc_set_switch) repair a differing step.
 *
 * Plain C99, one environment per OrcSim, scalar double arithmetic.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lcr_model.h"

#define MINVAL 1e-15
#define MINIMP 1e-4
#define MAXIMP 0.9999
#define MINMU 1e-5
#define MAXVAL 1e10
#define NB (LCR_NABODY + LCR_MAXBOX) /* pose slots: 7 arm bodies, then the boxes (cubes, then the static walls) */
#define NV LCR_MAXNV

typedef struct {
  double pos[3], frame[9], dist, friction[5], solref[2], solimp[5], mu;
  int dim, b1, b2, g1, g2, efc;
} Contact;

typedef struct OrcSim {
  LcrModel m;
  LcrEnvCfg cfg;
  double *verts;
  /* persistent state (MjData fields the reference carries across steps) */
  double qpos[LCR_MAXNQ], qvel[NV], ctrl[LCR_NARM], warm[NV], time;
  double target[3];
  double site_xpos[3], cube_xpos[LCR_MAXCUBE][3]; /* as of the last kinematics pass (stale-read rule) */
  int elapsed, needs_reset;
  uint64_t rng[4]; /* PCG64: state_hi, state_lo, inc_hi, inc_lo */
  /* per-forward data */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9], axis[LCR_NARM][3];
  double M[NV][NV], Lm[NV][NV];
  double bias[NV], passive[NV], actuator[NV], smooth[NV], qacc_smooth[NV], qacc[NV], qfrc_constraint[NV];
  Contact con[LCR_MAXCON_BIG];
  int ncon, nefc, niter, overflow, nan_resets, max_nefc;
  int peak_ncon, peak_nefc, peak_ncand, total_overflow; /* over the lifetime of the sim (test statistics) */
  /* Named switches for the choices that could not be checked against MuJoCo's source ("MJ choices"): a
   * mismatch against a MuJoCo fixture (tests/test_golden_mujoco.py) can be bisected by flipping them one at a time.
   *   plane_hull_tilt           tilt of the 3 extra support directions of plane-vs-hull contacts (default 1e-3)
   *   implicit_kv_when_clamped  1 (default): -kv stays in the implicitfast derivative when the servo force is clamped;
   *                             0: a dof whose actuator force sits at actuatorfrcrange contributes only its damping
   *   impratio                  > 0 overrides the compiled model's value (the <option> merge of the include chain) */
  double sw_tilt, sw_impratio;
  int sw_kv_clamped;
  int hist_nefc[32], hist_ncon[32]; /* env.steps by their max nefc / 16 and max ncon / 8 */
  int step_max_ncon;
  int sa_key[LCR_NSA], sa_next; /* separating-axis cache, see lcr_oracle_convex.inc */
  double sa_dir[LCR_NSA][3], sa_S[LCR_NSA][2], sa_u[LCR_NSA][2][3];
  int efc_type[LCR_MAXEFC_BIG]; /* 0 limit, 1 first row of a contact, 2 following row of a contact */
  int efc_con[LCR_MAXEFC_BIG];
  double J[LCR_MAXEFC_BIG][NV], efc_pos[LCR_MAXEFC_BIG], efc_vel[LCR_MAXEFC_BIG], efc_diag[LCR_MAXEFC_BIG];
  double efc_R[LCR_MAXEFC_BIG], efc_D[LCR_MAXEFC_BIG], efc_aref[LCR_MAXEFC_BIG], efc_force[LCR_MAXEFC_BIG], efc_jar[LCR_MAXEFC_BIG];
} OrcSim;

/* ------------------------------------------------------------------ small math */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
static void quat_mul(double *r, const double *a, const double *b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_to_mat(double *m, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void quat_normalize(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double inv = 1 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
static void mat_vec(double *r, const double *m, const double *v) { /* r = m v, row-major 3x3 */
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void matT_vec(double *r, const double *m, const double *v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------ PCG64 (numpy bit generator)
 * gymnasium seeds Env.np_random = numpy Generator(PCG64(SeedSequence(seed))); the reference draws
 * reset positions with np_random.uniform(low, high) (reach_cube_env.py:302).  numpy's PCG64 is the
 * 128-bit LCG "setseq" variant with the XSL-RR 64-bit output; Generator.uniform is
 * low + (high-low) * ((next64 >> 11) * 2^-53), one draw per array element. */
typedef unsigned __int128 u128;
static uint64_t pcg64_next(uint64_t s[4]) {
  const u128 mult = (((u128)0x2360ED051FC65DA4ULL) << 64) | 0x4385DF649FCCF645ULL;
  u128 state = (((u128)s[0]) << 64) | s[1], inc = (((u128)s[2]) << 64) | s[3];
  state = state * mult + inc;
  s[0] = (uint64_t)(state >> 64); s[1] = (uint64_t)state;
  uint64_t x = s[0] ^ s[1];
  unsigned rot = (unsigned)(s[0] >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
static double pcg64_double(uint64_t s[4]) { return (double)(pcg64_next(s) >> 11) * (1.0 / 9007199254740992.0); }
static void draw_uniform3(uint64_t s[4], const double *lo, const double *hi, double *out) {
  for (int k = 0; k < 3; k++) out[k] = lo[k] + (hi[k] - lo[k]) * pcg64_double(s);
}

/* ------------------------------------------------------------------ kinematics (mj_kinematics, mj_comPos) */
static void kinematics(OrcSim *s) {
  const LcrModel *m = &s->m;
  double p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0}, R[9], t[3], ql[4];
  for (int b = 0; b < LCR_NABODY; b++) {
    quat_to_mat(R, q);
    mat_vec(t, R, m->body_pos[b]);
    p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
    quat_mul(q, q, m->body_quat[b]);
    if (b >= 1) {
      int j = b - 1;
      quat_to_mat(R, q);
      mat_vec(s->axis[j], R, m->jnt_axis[j]);
      double h = 0.5 * s->qpos[j], sn = sin(h);
      ql[0] = cos(h); ql[1] = sn * m->jnt_axis[j][0]; ql[2] = sn * m->jnt_axis[j][1]; ql[3] = sn * m->jnt_axis[j][2];
      quat_mul(q, q, ql);
    }
    memcpy(s->xpos[b], p, sizeof p);
    memcpy(s->xquat[b], q, sizeof q);
    quat_to_mat(s->xmat[b], q);
    mat_vec(t, s->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) s->xipos[b][k] = p[k] + t[k];
    quat_mul(ql, q, m->body_iquat[b]);
    quat_to_mat(s->ximat[b], ql);
  }
  for (int c = 0; c < m->ncube; c++) {
    int b = LCR_NABODY + c;
    const double *qp = s->qpos + LCR_NARM + 7 * c;
    memcpy(s->xpos[b], qp, 3 * sizeof(double));
    memcpy(s->xquat[b], qp + 3, 4 * sizeof(double));
    quat_normalize(s->xquat[b]);
    quat_to_mat(s->xmat[b], s->xquat[b]);
    memcpy(s->xipos[b], s->xpos[b], 3 * sizeof(double));
    memcpy(s->ximat[b], s->xmat[b], 9 * sizeof(double));
    memcpy(s->cube_xpos[c], s->xpos[b], 3 * sizeof(double));
  }
  for (int w = 0; w < m->nwall; w++) { /* static boxes of the world body: constant pose in the slots after the cubes */
    int b = LCR_NABODY + m->ncube + w;
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, q1[4] = {1, 0, 0, 0};
    memcpy(s->xpos[b], m->wall_pos[w], 3 * sizeof(double));
    memcpy(s->xquat[b], q1, sizeof q1);
    memcpy(s->xmat[b], I3, sizeof I3);
  }
  mat_vec(t, s->xmat[m->site_body], m->site_pos);
  for (int k = 0; k < 3; k++) s->site_xpos[k] = s->xpos[m->site_body][k] + t[k];
}

/* translational (jp) and rotational (jr) Jacobian columns of a point fixed to body b (3 x nv) */
static void jac_point(const OrcSim *s, int b, const double *pt, double jp[3][NV], double jr[3][NV]) {
  memset(jp, 0, sizeof(double) * 3 * NV);
  memset(jr, 0, sizeof(double) * 3 * NV);
  if (b < 0) return;
  if (b < LCR_NABODY) {
    for (int j = 0; j < b; j++) { /* joint j sits on body j+1 <= b */
      double r[3], c[3];
      sub3(r, pt, s->xpos[j + 1]);
      cross3(c, s->axis[j], r);
      for (int k = 0; k < 3; k++) { jp[k][j] = c[k]; jr[k][j] = s->axis[j][k]; }
    }
  } else {
    int c = b - LCR_NABODY, d0 = LCR_NARM + 6 * c;
    double r[3];
    sub3(r, pt, s->xpos[b]);
    for (int k = 0; k < 3; k++) {
      jp[k][d0 + k] = 1.0;
      double ax[3] = {s->xmat[b][k], s->xmat[b][3 + k], s->xmat[b][6 + k]}, cr[3]; /* body axis k in world */
      cross3(cr, ax, r);
      for (int i = 0; i < 3; i++) { jp[i][d0 + 3 + k] = cr[i]; jr[i][d0 + 3 + k] = ax[i]; }
    }
  }
}

/* ------------------------------------------------------------------ inertia + bias (mj_crb, mj_rne) */
static void mass_matrix(OrcSim *s) {
  const LcrModel *m = &s->m;
  int nv = m->nv;
  memset(s->M, 0, sizeof s->M);
  double jp[3][NV], jr[3][NV];
  for (int b = 1; b < LCR_NABODY; b++) {
    jac_point(s, b, s->xipos[b], jp, jr);
    double Iw[9]; /* ximat diag(I) ximat^T */
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) {
        double a = 0;
        for (int d = 0; d < 3; d++) a += s->ximat[b][3 * i + d] * m->body_inertia[b][d] * s->ximat[b][3 * k + d];
        Iw[3 * i + k] = a;
      }
    for (int i = 0; i < LCR_NARM; i++)
      for (int j = 0; j < LCR_NARM; j++) {
        double a = 0;
        for (int k = 0; k < 3; k++) a += m->body_mass[b] * jp[k][i] * jp[k][j];
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) a += jr[k][i] * Iw[3 * k + l] * jr[l][j];
        s->M[i][j] += a;
      }
  }
  for (int j = 0; j < LCR_NARM; j++) s->M[j][j] += m->jnt_armature[j];
  for (int c = 0; c < m->ncube; c++)
    for (int k = 0; k < 3; k++) {
      s->M[LCR_NARM + 6 * c + k][LCR_NARM + 6 * c + k] = m->cube_mass[c];
      s->M[LCR_NARM + 6 * c + 3 + k][LCR_NARM + 6 * c + 3 + k] = m->cube_inertia[c][k];
    }
  (void)nv;
}

static void bias_forces(OrcSim *s) {
  const LcrModel *m = &s->m;
  double w[LCR_NABODY][3], al[LCR_NABODY][3], a[LCR_NABODY][3], F[LCR_NABODY][3], N[LCR_NABODY][3];
  memset(w, 0, sizeof w); memset(al, 0, sizeof al);
  for (int k = 0; k < 3; k++) a[0][k] = -m->gravity[k]; /* gravity as base acceleration */
  memset(F, 0, sizeof F); memset(N, 0, sizeof N);
  for (int b = 1; b < LCR_NABODY; b++) {
    int j = b - 1;
    double zq[3], r[3], t1[3], t2[3];
    for (int k = 0; k < 3; k++) zq[k] = s->axis[j][k] * s->qvel[j];
    cross3(t1, w[b - 1], zq);
    for (int k = 0; k < 3; k++) { w[b][k] = w[b - 1][k] + zq[k]; al[b][k] = al[b - 1][k] + t1[k]; }
    sub3(r, s->xpos[b], s->xpos[b - 1]);
    cross3(t1, al[b - 1], r);
    cross3(t2, w[b - 1], r);
    cross3(t2, w[b - 1], t2);
    for (int k = 0; k < 3; k++) a[b][k] = a[b - 1][k] + t1[k] + t2[k];
    double rc[3], ac[3];
    sub3(rc, s->xipos[b], s->xpos[b]);
    cross3(t1, al[b], rc);
    cross3(t2, w[b], rc);
    cross3(t2, w[b], t2);
    for (int k = 0; k < 3; k++) ac[k] = a[b][k] + t1[k] + t2[k];
    for (int k = 0; k < 3; k++) F[b][k] = m->body_mass[b] * ac[k];
    /* N = Iw al + w x (Iw w), Iw = R diag(I) R^T */
    double lw[3], la[3], Il[3], Ia[3], Iww[3], Iwa[3];
    matT_vec(lw, s->ximat[b], w[b]);
    matT_vec(la, s->ximat[b], al[b]);
    for (int k = 0; k < 3; k++) { Il[k] = m->body_inertia[b][k] * lw[k]; Ia[k] = m->body_inertia[b][k] * la[k]; }
    mat_vec(Iww, s->ximat[b], Il);
    mat_vec(Iwa, s->ximat[b], Ia);
    cross3(t1, w[b], Iww);
    for (int k = 0; k < 3; k++) N[b][k] = Iwa[k] + t1[k];
  }
  memset(s->bias, 0, sizeof s->bias);
  for (int j = 0; j < LCR_NARM; j++) {
    double tq[3] = {0, 0, 0};
    for (int b = j + 1; b < LCR_NABODY; b++) {
      double r[3], t[3];
      sub3(r, s->xipos[b], s->xpos[j + 1]);
      cross3(t, r, F[b]);
      for (int k = 0; k < 3; k++) tq[k] += N[b][k] + t[k];
    }
    s->bias[j] = dot3(s->axis[j], tq);
  }
  for (int c = 0; c < m->ncube; c++) /* isotropic inertia: w x Iw = 0 */
    for (int k = 0; k < 3; k++) s->bias[LCR_NARM + 6 * c + k] = -m->cube_mass[c] * m->gravity[k];
}

/* dense Cholesky of the leading n x n block: A = L L^T; returns 0 on success */
static int cholesky(int n, double A[NV][NV], double L[NV][NV]) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double a = A[i][j];
      for (int k = 0; k < j; k++) a -= L[i][k] * L[j][k];
      if (i == j) {
        if (a < MINVAL) a = MINVAL;
        L[i][i] = sqrt(a);
      } else
        L[i][j] = a / L[j][j];
    }
  return 0;
}
static void chol_solve(int n, double L[NV][NV], const double *b, double *x) {
  double y[NV];
  for (int i = 0; i < n; i++) {
    double a = b[i];
    for (int k = 0; k < i; k++) a -= L[i][k] * y[k];
    y[i] = a / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double a = y[i];
    for (int k = i + 1; k < n; k++) a -= L[k][i] * x[k];
    x[i] = a / L[i][i];
  }
}

/* passive + actuator + smooth acceleration (mj_passive, mj_fwdActuation, mj_fwdAcceleration) */
static void smooth_forces(OrcSim *s) {
  const LcrModel *m = &s->m;
  int nv = m->nv;
  memset(s->passive, 0, sizeof s->passive);
  memset(s->actuator, 0, sizeof s->actuator);
  for (int j = 0; j < LCR_NARM; j++) {
    s->passive[j] = -m->jnt_damping[j] * s->qvel[j];
    double u = clampd(s->ctrl[j], m->act_ctrlrange[j][0], m->act_ctrlrange[j][1]);
    double f = m->act_kp[j] * (u - s->qpos[j]) - m->act_kv[j] * s->qvel[j];
    s->actuator[j] = clampd(f, m->jnt_frcrange[j][0], m->jnt_frcrange[j][1]);
  }
  for (int i = 0; i < nv; i++) s->smooth[i] = s->passive[i] - s->bias[i] + s->actuator[i];
  cholesky(nv, s->M, s->Lm);
  chol_solve(nv, s->Lm, s->smooth, s->qacc_smooth);
}

/* ------------------------------------------------------------------ collision */
static void make_frame(double *f) { /* f[0:3] = unit normal; builds the two tangents (mju_makeFrame) */
  double y[3] = {0, 0, 0};
  if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
  double d = dot3(f, y);
  for (int k = 0; k < 3; k++) y[k] -= d * f[k];
  double inv = 1 / sqrt(dot3(y, y));
  for (int k = 0; k < 3; k++) f[3 + k] = y[k] * inv;
  cross3(f + 6, f, f + 3);
}

/* contact parameter mixing of two geoms (MuJoCo: priority, then max condim / max friction / solmix average) */
static void mix_params(const LcrModel *m, int g1, int g2, Contact *c) {
  double fr[3];
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  if (p1 != p2) {
    int g = p1 > p2 ? g1 : g2;
    c->dim = m->geom_condim[g];
    memcpy(fr, m->geom_friction[g], sizeof fr);
    memcpy(c->solref, m->geom_solref[g], sizeof c->solref);
    memcpy(c->solimp, m->geom_solimp[g], sizeof c->solimp);
  } else {
    c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    for (int k = 0; k < 3; k++) fr[k] = fmax(m->geom_friction[g1][k], m->geom_friction[g2][k]);
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else mix = s1 < MINVAL ? 0.0 : 1.0;
    if (m->geom_solref[g1][0] > 0 && m->geom_solref[g2][0] > 0)
      for (int k = 0; k < 2; k++) c->solref[k] = mix * m->geom_solref[g1][k] + (1 - mix) * m->geom_solref[g2][k];
    else
      for (int k = 0; k < 2; k++) c->solref[k] = fmin(m->geom_solref[g1][k], m->geom_solref[g2][k]);
    for (int k = 0; k < 5; k++) c->solimp[k] = mix * m->geom_solimp[g1][k] + (1 - mix) * m->geom_solimp[g2][k];
  }
  c->friction[0] = c->friction[1] = fmax(MINMU, fr[0]);
  c->friction[2] = fmax(MINMU, fr[1]);
  c->friction[3] = c->friction[4] = fmax(MINMU, fr[2]);
  c->g1 = g1; c->g2 = g2;
}

/* boxes: index c < ncube = free cube c, c >= ncube = static wall c - ncube; pose slot LCR_NABODY + c, geom nmesh + 1 + c */
static const double *box_half(const LcrModel *m, int c) { return c < m->ncube ? m->cube_size[c] : m->wall_size[c - m->ncube]; }
static int geom_body(const LcrModel *m, int g) {
  if (g < m->nmesh) return m->mesh_body[g];
  if (g == m->nmesh) return -1;
  return g - m->nmesh - 1 < m->ncube ? LCR_NABODY + (g - m->nmesh - 1) : -1; /* walls belong to the world */
}

static Contact *add_contact(OrcSim *s, int g1, int g2, const double *pos, const double *normal, double dist) {
  if (s->ncon >= LCR_MAXCON_BIG) { s->overflow++; return NULL; }
  Contact *c = &s->con[s->ncon];
  mix_params(&s->m, g1, g2, c);
  if (s->nefc + c->dim > LCR_MAXEFC_BIG) { s->overflow++; return NULL; }
  memcpy(c->pos, pos, 3 * sizeof(double));
  memcpy(c->frame, normal, 3 * sizeof(double));
  make_frame(c->frame);
  c->dist = dist;
  c->b1 = geom_body(&s->m, g1);
  c->b2 = geom_body(&s->m, g2);
  c->efc = s->nefc;
  s->nefc += c->dim; /* rows are filled in make_constraints; reserved here so the cap is exact */
  s->ncon++;
  return c;
}

/* floor plane (z = 0, normal +z) vs box: one contact per corner below the plane, at most 4
 * (restates MuJoCo mjc_PlaneBox: corner loop in bit order, skip corners on the far side of the centre) */
static void collide_floor_cube(OrcSim *s, int c) {
  const LcrModel *m = &s->m;
  int b = LCR_NABODY + c, cnt = 0;
  const double n[3] = {0, 0, 1};
  double dist = s->xpos[b][2];
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double v[3] = {m->cube_size[c][0] * ((i & 1) ? 1 : -1), m->cube_size[c][1] * ((i & 2) ? 1 : -1),
                   m->cube_size[c][2] * ((i & 4) ? 1 : -1)}, w[3];
    mat_vec(w, s->xmat[b], v);
    double ld = w[2];
    if (dist + ld > 0 || ld > 0) continue;
    double d = dist + ld, pos[3] = {s->xpos[b][0] + w[0], s->xpos[b][1] + w[1], s->xpos[b][2] + w[2] - 0.5 * d};
    if (d >= 0) continue; /* margin 0: contacts exist only when penetrating */
    if (add_contact(s, m->nmesh, m->nmesh + 1 + c, pos, n, d)) cnt++;
  }
}

/* support vertex of mesh g along world direction d: returns the pool index, writes the world point */
static int mesh_support(const OrcSim *s, int g, const double *d, double *out) {
  const LcrModel *m = &s->m;
  int b = m->mesh_body[g];
  double dl[3];
  matT_vec(dl, s->xmat[b], d);
  int best = -1;
  double bv = -1e300;
  for (int i = 0; i < m->mesh_vertnum[g]; i++) {
    const double *v = s->verts + 3 * (m->mesh_vertadr[g] + i);
    double t = dot3(v, dl);
    if (t > bv) { bv = t; best = i; }
  }
  const double *v = s->verts + 3 * (m->mesh_vertadr[g] + best);
  mat_vec(out, s->xmat[b], v);
  for (int k = 0; k < 3; k++) out[k] += s->xpos[b][k];
  return best;
}

/* floor plane vs convex mesh: deepest vertex + up to 3 more support vertices in directions tilted
 * 120 degrees apart about -n (restates the scheme of MuJoCo mjc_PlaneConvex; tilt is ours: 1e-3) */
static void collide_floor_mesh(OrcSim *s, int g) {
  const LcrModel *m = &s->m;
  int b = m->mesh_body[g];
  double c[3];
  mat_vec(c, s->xmat[b], m->mesh_center[g]);
  if (s->xpos[b][2] + c[2] - m->mesh_rbound[g] > 0) return; /* bounding sphere above the floor */
  const double n[3] = {0, 0, 1};
  const double tl = s->sw_tilt;
  const double dirs[4][3] = {{0, 0, -1}, {tl, 0, -1}, {-0.5 * tl, 0.8660254037844386 * tl, -1}, {-0.5 * tl, -0.8660254037844386 * tl, -1}};
  int used[4], cnt = 0;
  for (int t = 0; t < 4; t++) {
    double p[3];
    int vi = mesh_support(s, g, dirs[t], p);
    double d = p[2];
    if (d >= 0) { if (t == 0) return; continue; }
    int dup = 0;
    for (int k = 0; k < cnt; k++) dup |= (used[k] == vi);
    if (dup) continue;
    double pos[3] = {p[0], p[1], p[2] - 0.5 * d};
    if (add_contact(s, m->nmesh, g, pos, n, d)) used[cnt++] = vi;
  }
}

#include "lcr_oracle_convex.inc"

static void collision(OrcSim *s) {
  const LcrModel *m = &s->m;
  int mask = s->cfg.collision_mask;
  s->ncon = 0;
  /* convex pairs: candidates -> jobs (against the cache of the substep start) -> consume, as in the CUDA kernels */
  int keys[LCR_MAXCAND];
  CandRes res[LCR_MAXCAND];
  const int ncand_all = collect_candidates(s, keys), ncand = ncand_all < LCR_MAXCAND ? ncand_all : LCR_MAXCAND;
  if (ncand_all > s->peak_ncand) s->peak_ncand = ncand_all;
  for (int k = 0; k < ncand; k++) cand_job(s, keys[k], &res[k]);
  /* generation order (= drop order at the caps): floor-cube, cube-cube, wall-cube, cube-mesh, wall-mesh, floor-mesh, mesh-mesh */
  if (mask & LCR_COLLIDE_FLOOR_CUBE)
    for (int c = 0; c < m->ncube; c++) collide_floor_cube(s, c);
  if ((mask & LCR_COLLIDE_CUBE_CUBE) && m->ncube == 2) collide_box_box(s, 0, 1);
  if (mask & LCR_COLLIDE_WALL_CUBE)
    for (int w = 0; w < m->nwall; w++)
      for (int c = 0; c < m->ncube; c++) collide_box_box(s, m->ncube + w, c);
  if (mask & (LCR_COLLIDE_CUBE_MESH | LCR_COLLIDE_WALL_MESH))
    for (int k = 0; k < ncand; k++) if (keys[k] >= 200) cand_apply(s, &res[k]);
  if (mask & LCR_COLLIDE_FLOOR_MESH)
    for (int g = 0; g < m->nmesh; g++)
      if (m->mesh_body[g] != 0) collide_floor_mesh(s, g);
  if (mask & LCR_COLLIDE_MESH_MESH) {
    for (int k = 0; k < ncand; k++) if (keys[k] < 200) cand_apply(s, &res[k]);
    if (ncand_all > LCR_MAXCAND) s->overflow += ncand_all - LCR_MAXCAND; /* candidates past the cap are not processed */
  }
}

/* ------------------------------------------------------------------ constraints (mj_makeConstraint, mj_makeImpedance, mj_referenceConstraint) */
static void sol_params(const double *solref, const double *solimp_in, double timestep, double pos, double *imp, double *K, double *B) {
  double si[5] = {clampd(solimp_in[0], MINIMP, MAXIMP), clampd(solimp_in[1], MINIMP, MAXIMP), fmax(0, solimp_in[2]),
                  clampd(solimp_in[3], MINIMP, MAXIMP), fmax(1, solimp_in[4])};
  /* impedance d(|pos|) */
  if (si[0] == si[1] || si[2] <= MINVAL) *imp = 0.5 * (si[0] + si[1]);
  else {
    double x = fabs(pos) / si[2];
    if (x >= 1) *imp = si[1];
    else if (x <= 0) *imp = si[0];
    else {
      double y;
      if (si[4] == 1) y = x;
      else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
      else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
      *imp = si[0] + y * (si[1] - si[0]);
    }
  }
  double dmax = si[1];
  if (solref[0] > 0) {
    double tc = fmax(solref[0], 2 * timestep), dr = solref[1]; /* refsafe */
    *K = 1 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
    *B = 2 / fmax(MINVAL, dmax * tc);
  } else {
    *K = -solref[0] / fmax(MINVAL, dmax * dmax);
    *B = -solref[1] / fmax(MINVAL, dmax);
  }
}

static double body_invweight(const LcrModel *m, int b, int rot) {
  if (b < 0) return 0;
  if (b < LCR_NABODY) return m->body_invweight0[b][rot];
  return m->cube_invweight0[b - LCR_NABODY][rot];
}

static void make_constraints(OrcSim *s) {
  const LcrModel *m = &s->m;
  int nv = m->nv;
  s->nefc = 0;
  s->overflow = 0;
  /* joint limits (margin 0): rows exist only beyond the range */
  for (int j = 0; j < LCR_NARM; j++)
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? s->qpos[j] - m->jnt_range[j][0] : m->jnt_range[j][1] - s->qpos[j];
      if (dist >= 0) continue;
      int i = s->nefc++;
      memset(s->J[i], 0, sizeof s->J[i]);
      s->J[i][j] = side == 0 ? 1.0 : -1.0;
      s->efc_type[i] = 0; s->efc_con[i] = j;
      s->efc_pos[i] = dist;
      s->efc_diag[i] = m->dof_invweight0[j];
    }
  collision(s);
  double jp1[3][NV], jr1[3][NV], jp2[3][NV], jr2[3][NV];
  for (int ci = 0; ci < s->ncon; ci++) {
    Contact *c = &s->con[ci];
    jac_point(s, c->b1, c->pos, jp1, jr1);
    jac_point(s, c->b2, c->pos, jp2, jr2);
    for (int r = 0; r < c->dim; r++) {
      int i = c->efc + r;
      const double *ax = c->frame + 3 * (r % 3);
      for (int d = 0; d < nv; d++) {
        double a = 0;
        if (r < 3) for (int k = 0; k < 3; k++) a += ax[k] * (jp2[k][d] - jp1[k][d]);
        else for (int k = 0; k < 3; k++) a += ax[k] * (jr2[k][d] - jr1[k][d]);
        s->J[i][d] = a;
      }
      for (int d = nv; d < NV; d++) s->J[i][d] = 0;
      s->efc_type[i] = r == 0 ? 1 : 2; s->efc_con[i] = ci;
      s->efc_pos[i] = r == 0 ? c->dist : 0.0;
      s->efc_diag[i] = body_invweight(m, c->b1, r >= 3) + body_invweight(m, c->b2, r >= 3);
    }
  }
  /* impedance, regularisation, reference acceleration */
  for (int i = 0; i < s->nefc; i++) {
    double v = 0;
    for (int d = 0; d < nv; d++) v += s->J[i][d] * s->qvel[d];
    s->efc_vel[i] = v;
    const double *solref, *solimp;
    if (s->efc_type[i] == 0) { solref = m->jnt_solref[s->efc_con[i]]; solimp = m->jnt_solimp[s->efc_con[i]]; }
    else { solref = s->con[s->efc_con[i]].solref; solimp = s->con[s->efc_con[i]].solimp; }
    double imp, K, B;
    sol_params(solref, solimp, m->timestep, s->efc_pos[i], &imp, &K, &B);
    s->efc_R[i] = fmax(MINVAL, (1 - imp) * s->efc_diag[i] / imp);
    s->efc_aref[i] = -B * v - K * imp * s->efc_pos[i];
  }
  for (int ci = 0; ci < s->ncon; ci++) { /* elliptic friction rows: R from the normal row and impratio */
    Contact *c = &s->con[ci];
    int i = c->efc;
    if (c->dim > 1) {
      s->efc_R[i + 1] = s->efc_R[i] / fmax(MINVAL, s->sw_impratio > 0 ? s->sw_impratio : m->impratio);
      for (int j = 1; j < c->dim; j++)
        s->efc_R[i + j] = s->efc_R[i + 1] * c->friction[0] * c->friction[0] / (c->friction[j - 1] * c->friction[j - 1]);
      c->mu = c->friction[0] * sqrt(s->efc_R[i + 1] / s->efc_R[i]);
    } else c->mu = 0;
  }
  for (int i = 0; i < s->nefc; i++) s->efc_D[i] = 1 / s->efc_R[i];
  if (s->nefc > s->max_nefc) s->max_nefc = s->nefc;
  if (s->nefc > s->peak_nefc) s->peak_nefc = s->nefc;
  if (s->ncon > s->peak_ncon) s->peak_ncon = s->ncon;
  if (s->ncon > s->step_max_ncon) s->step_max_ncon = s->ncon;
  s->total_overflow += s->overflow;
}

/* ------------------------------------------------------------------ primal Newton solver (mj_solNewton) */
/* cost, first and second derivative along jar + alpha*jv of one constraint unit (limit row or
 * elliptic contact); if force != NULL also writes efc_force, and if Hc != NULL (contact in the
 * middle zone) the dim x dim cone Hessian in jar space.  Returns the zone: 0 satisfied, 1 quadratic,
 * 2 cone (middle). */
static int unit_eval(const OrcSim *s, int i, const double *jar, const double *jv, double alpha, double *cost, double *d1,
                     double *d2, double *force, double *Hc) {
  *cost = *d1 = *d2 = 0;
  if (s->efc_type[i] == 0) {
    double x = jar[i] + (jv ? alpha * jv[i] : 0);
    if (x < 0) {
      *cost = 0.5 * s->efc_D[i] * x * x;
      if (jv) { *d1 = s->efc_D[i] * x * jv[i]; *d2 = s->efc_D[i] * jv[i] * jv[i]; }
      if (force) force[i] = -s->efc_D[i] * x;
      return 1;
    }
    if (force) force[i] = 0;
    return 0;
  }
  const Contact *c = &s->con[s->efc_con[i]];
  int dim = c->dim;
  double x[6], u[6] = {0, 0, 0, 0, 0, 0}, fri[6], mu = c->mu;
  for (int j = 0; j < dim; j++) x[j] = jar[i + j] + (jv ? alpha * jv[i + j] : 0);
  if (dim == 1) {
    if (x[0] < 0) {
      *cost = 0.5 * s->efc_D[i] * x[0] * x[0];
      if (jv) { *d1 = s->efc_D[i] * x[0] * jv[i]; *d2 = s->efc_D[i] * jv[i] * jv[i]; }
      if (force) force[i] = -s->efc_D[i] * x[0];
      return 1;
    }
    if (force) force[i] = 0;
    return 0;
  }
  fri[0] = mu;
  for (int j = 1; j < dim; j++) fri[j] = c->friction[j - 1];
  double T2 = 0;
  for (int j = 0; j < dim; j++) { u[j] = x[j] * fri[j]; if (j) T2 += u[j] * u[j]; }
  double N = u[0], T = sqrt(T2);
  if (N >= mu * T || (T <= 0 && N >= 0)) { /* top zone */
    if (force) for (int j = 0; j < dim; j++) force[i + j] = 0;
    return 0;
  }
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) { /* bottom zone: all rows quadratic */
    for (int j = 0; j < dim; j++) {
      *cost += 0.5 * s->efc_D[i + j] * x[j] * x[j];
      if (jv) { *d1 += s->efc_D[i + j] * x[j] * jv[i + j]; *d2 += s->efc_D[i + j] * jv[i + j] * jv[i + j]; }
      if (force) force[i + j] = -s->efc_D[i + j] * x[j];
    }
    return 1;
  }
  /* middle zone */
  double Dm = s->efc_D[i] / (mu * mu * (1 + mu * mu)), NT = N - mu * T;
  *cost = 0.5 * Dm * NT * NT;
  if (jv) {
    double N1 = mu * jv[i], T1 = 0, up2 = 0;
    for (int j = 1; j < dim; j++) { double up = fri[j] * jv[i + j]; T1 += u[j] * up; up2 += up * up; }
    T1 /= T;
    double T2d = up2 / T - T1 * T1 / T;
    double NT1 = N1 - mu * T1, NT2 = -mu * T2d;
    *d1 = Dm * NT * NT1;
    *d2 = Dm * (NT1 * NT1 + NT * NT2);
  }
  if (force) {
    force[i] = -Dm * NT * mu;
    for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * u[j] * fri[j];
  }
  if (Hc) {
    double g[6];
    g[0] = mu;
    for (int j = 1; j < dim; j++) g[j] = -mu * fri[j] * u[j] / T;
    for (int a = 0; a < dim; a++)
      for (int b = 0; b < dim; b++) {
        double h = 0;
        if (a && b) h = -mu * fri[a] * fri[b] * ((a == b ? 1.0 / T : 0.0) - u[a] * u[b] / (T * T * T));
        Hc[a * 6 + b] = Dm * (g[a] * g[b] + NT * h);
      }
  }
  return 2;
}

static int unit_rows(const OrcSim *s, int i) { return s->efc_type[i] == 0 ? 1 : s->con[s->efc_con[i]].dim; }

/* total cost at qacc (Gauss + constraints); fills jar, Ma; optionally force and gradient */
static double total_cost(OrcSim *s, const double *qacc, double *jar, double *Ma, double *force, double *grad) {
  int nv = s->m.nv;
  for (int i = 0; i < s->nefc; i++) {
    double a = -s->efc_aref[i];
    for (int d = 0; d < nv; d++) a += s->J[i][d] * qacc[d];
    jar[i] = a;
  }
  double cost = 0;
  for (int i = 0; i < nv; i++) {
    double a = 0;
    for (int d = 0; d < nv; d++) a += s->M[i][d] * qacc[d];
    Ma[i] = a;
    cost += 0.5 * (a - s->smooth[i]) * (qacc[i] - s->qacc_smooth[i]);
  }
  for (int i = 0; i < s->nefc; i += unit_rows(s, i)) {
    double c, d1, d2;
    unit_eval(s, i, jar, NULL, 0, &c, &d1, &d2, force, NULL);
    cost += c;
  }
  if (grad && force)
    for (int d = 0; d < nv; d++) {
      double a = Ma[d] - s->smooth[d];
      for (int i = 0; i < s->nefc; i++) a -= s->J[i][d] * force[i];
      grad[d] = a;
    }
  return cost;
}

static void solve_constraints(OrcSim *s) {
  const LcrModel *m = &s->m;
  int nv = m->nv, nefc = s->nefc;
  s->niter = 0;
  if (nefc == 0) {
    memcpy(s->qacc, s->qacc_smooth, sizeof s->qacc);
    memcpy(s->warm, s->qacc_smooth, sizeof s->warm);
    memset(s->qfrc_constraint, 0, sizeof s->qfrc_constraint);
    return;
  }
  double jar[LCR_MAXEFC_BIG], Ma[NV], grad[NV], search[NV], Mv[NV], jv[LCR_MAXEFC_BIG];
  double H[NV][NV], L[NV][NV];
  double *force = s->efc_force, *qacc = s->qacc;
  /* if no constraint row touches an arm dof (no limit rows, no contact on links 1..6) the arm block is decoupled:
   * those dofs take no warm start, so they stay at qacc_smooth (the optimum) through the monolithic iterations */
  {
    int arm = 0;
    for (int i = 0; i < nefc; i++) if (s->efc_type[i] == 0) arm = 1;
    for (int c = 0; c < s->ncon; c++) if ((s->con[c].b1 > 0 && s->con[c].b1 < LCR_NABODY) || (s->con[c].b2 > 0 && s->con[c].b2 < LCR_NABODY)) arm = 1;
    if (!arm) memcpy(s->warm, s->qacc_smooth, sizeof(double) * LCR_NARM);
  }
  /* warm start: the better of qacc_warmstart and qacc_smooth */
  double cw = total_cost(s, s->warm, jar, Ma, NULL, NULL);
  double cs = total_cost(s, s->qacc_smooth, jar, Ma, NULL, NULL);
  memcpy(qacc, cw < cs ? s->warm : s->qacc_smooth, sizeof(double) * NV);
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double cost = total_cost(s, qacc, jar, Ma, force, grad);
  for (int iter = 0; iter < m->iterations; iter++) {
    /* Hessian H = M + J^T D J over quadratic rows + cone blocks */
    for (int a = 0; a < nv; a++) for (int b = 0; b < nv; b++) H[a][b] = s->M[a][b];
    for (int i = 0; i < nefc;) {
      int n = unit_rows(s, i);
      double c, d1, d2, Hc[36];
      int zone = unit_eval(s, i, jar, NULL, 0, &c, &d1, &d2, NULL, Hc);
      if (zone == 1)
        for (int r = 0; r < n; r++)
          for (int a = 0; a < nv; a++) {
            double t = s->efc_D[i + r] * s->J[i + r][a];
            if (t != 0) for (int b = 0; b < nv; b++) H[a][b] += t * s->J[i + r][b];
          }
      else if (zone == 2)
        for (int r = 0; r < n; r++)
          for (int q = 0; q < n; q++) {
            double h = Hc[r * 6 + q];
            for (int a = 0; a < nv; a++) {
              double t = h * s->J[i + r][a];
              if (t != 0) for (int b = 0; b < nv; b++) H[a][b] += t * s->J[i + q][b];
            }
          }
      i += n;
    }
    cholesky(nv, H, L);
    chol_solve(nv, L, grad, search);
    double snorm = 0, gs = 0;
    for (int d = 0; d < nv; d++) { search[d] = -search[d]; snorm += search[d] * search[d]; gs += search[d] * grad[d]; }
    snorm = sqrt(snorm);
    if (snorm < MINVAL) break;
    /* exact line search on f(alpha) = cost(qacc + alpha*search): safeguarded Newton */
    for (int i = 0; i < nv; i++) { double a = 0; for (int d = 0; d < nv; d++) a += s->M[i][d] * search[d]; Mv[i] = a; }
    for (int i = 0; i < nefc; i++) { double a = 0; for (int d = 0; d < nv; d++) a += s->J[i][d] * search[d]; jv[i] = a; }
    double g1 = 0, g2 = 0; /* Gauss part: derivative at 0 and curvature */
    for (int d = 0; d < nv; d++) { g1 += search[d] * (Ma[d] - s->smooth[d]); g2 += search[d] * Mv[d]; }
    double gtol = m->tolerance * m->ls_tolerance * snorm / scale;
    double alpha = 0, lo = 0, hi = -1, d1 = 0, d2 = 0;
    for (int ls = 0; ls <= m->ls_iterations; ls++) {
      d1 = g1 + alpha * g2; d2 = g2;
      for (int i = 0; i < nefc; i += unit_rows(s, i)) {
        double c, a1, a2;
        unit_eval(s, i, jar, jv, alpha, &c, &a1, &a2, NULL, NULL);
        d1 += a1; d2 += a2;
      }
      if (fabs(d1) < gtol || ls == m->ls_iterations) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      double an = alpha - d1 / d2;
      if (hi >= 0 && (an <= lo || an >= hi)) an = 0.5 * (lo + hi);
      if (an == alpha) break;
      alpha = an;
    }
    if (alpha <= 0) break;
    for (int d = 0; d < nv; d++) qacc[d] += alpha * search[d];
    double old = cost;
    cost = total_cost(s, qacc, jar, Ma, force, grad);
    s->niter = iter + 1;
    double gn = 0;
    for (int d = 0; d < nv; d++) gn += grad[d] * grad[d];
    if (scale * (old - cost) < m->tolerance || scale * sqrt(gn) < m->tolerance) break;
  }
  for (int d = 0; d < nv; d++) {
    double a = 0;
    for (int i = 0; i < nefc; i++) a += s->J[i][d] * force[i];
    s->qfrc_constraint[d] = a;
  }
  memcpy(s->warm, qacc, sizeof s->warm);
}

/* ------------------------------------------------------------------ mj_forward / mj_step */
static void reset_data(OrcSim *s);

static void sa_clear(OrcSim *s) { for (int k = 0; k < LCR_NSA; k++) s->sa_key[k] = -1; s->sa_next = 0; }

static void forward_(OrcSim *s) {
  kinematics(s);
  mass_matrix(s);
  bias_forces(s);
  make_constraints(s);
  smooth_forces(s);
  solve_constraints(s);
}

static int bad(double x) { return !(x == x) || x > MAXVAL || x < -MAXVAL; }

static void substep_(OrcSim *s) {
  const LcrModel *m = &s->m;
  int nv = m->nv;
  double h = m->timestep;
  for (int i = 0; i < m->nq; i++) if (bad(s->qpos[i])) { reset_data(s); s->nan_resets++; break; }
  for (int i = 0; i < nv; i++) if (bad(s->qvel[i])) { reset_data(s); s->nan_resets++; break; }
  forward_(s);
  for (int i = 0; i < nv; i++) if (bad(s->qacc[i])) { reset_data(s); s->nan_resets++; forward_(s); break; }
  /* implicitfast: (M - h dF/dv) a = M qacc, dF/dv = -(damping + kv) on the arm dofs */
  double rhs[NV], a[NV];
  double A[NV][NV], L[NV][NV];
  for (int i = 0; i < nv; i++) {
    double t = 0;
    for (int d = 0; d < nv; d++) { t += s->M[i][d] * s->qacc[d]; A[i][d] = s->M[i][d]; }
    rhs[i] = t;
  }
  for (int j = 0; j < LCR_NARM; j++) {
    const int clamped = s->actuator[j] <= m->jnt_frcrange[j][0] || s->actuator[j] >= m->jnt_frcrange[j][1];
    A[j][j] += h * (m->jnt_damping[j] + ((s->sw_kv_clamped || !clamped) ? m->act_kv[j] : 0.0));
  }
  cholesky(nv, A, L);
  chol_solve(nv, L, rhs, a);
  for (int i = 0; i < nv; i++) s->qvel[i] += h * a[i];
  for (int j = 0; j < LCR_NARM; j++) s->qpos[j] += h * s->qvel[j];
  for (int c = 0; c < m->ncube; c++) {
    double *qp = s->qpos + LCR_NARM + 7 * c, *qv = s->qvel + LCR_NARM + 6 * c;
    for (int k = 0; k < 3; k++) qp[k] += h * qv[k];
    double w = norm3(qv + 3);
    quat_normalize(qp + 3);
    if (w * h > 0) { /* quat <- quat * exp(h w / 2), w body-local */
      double ang = w * h, sn = sin(0.5 * ang) / w, dq[4] = {cos(0.5 * ang), sn * qv[3], sn * qv[4], sn * qv[5]}, r[4];
      quat_mul(r, qp + 3, dq);
      memcpy(qp + 3, r, sizeof r);
      quat_normalize(qp + 3);
    }
  }
  s->time += h;
}

/* API entry points: the separating-axis cache lives for one call */
void orc_forward(OrcSim *s) { forward_(s); }
void orc_substeps(OrcSim *s, int n) { s->max_nefc = 0; for (int k = 0; k < n; k++) substep_(s); }
void orc_substep(OrcSim *s) { orc_substeps(s, 1); }

static void reset_data(OrcSim *s) { /* mj_resetData: qpos0, everything else zero */
  const LcrModel *m = &s->m;
  memset(s->qpos, 0, sizeof s->qpos); memset(s->qvel, 0, sizeof s->qvel);
  memset(s->ctrl, 0, sizeof s->ctrl); memset(s->warm, 0, sizeof s->warm);
  s->time = 0;
  for (int c = 0; c < m->ncube; c++) {
    memcpy(s->qpos + LCR_NARM + 7 * c, m->cube_pos0[c], 3 * sizeof(double));
    s->qpos[LCR_NARM + 7 * c + 3] = 1;
  }
}

/* ------------------------------------------------------------------ env glue */
static const double TARGET_LOW[6] = {-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533};
static const double TARGET_HIGH[6] = {3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599};

int orc_obs_dim(int task) { return (task == LCR_TASK_REACH || task == LCR_TASK_LIFT || task == LCR_TASK_PUSH_LOOP) ? 15 : 18; }
int orc_action_dim(const LcrEnvCfg *cfg) { return (cfg->action_mode ? 3 : 5) + (cfg->block_gripper ? 0 : 1); }

/* get_observation (reach_cube_env.py:281-295, push_cube_env.py:291-306, stack_two_cubes_env.py:290-305,
 * push_cube_loop_env.py:286-300: arm_qpos, arm_qvel, cube_pos like Reach) */
static void write_obs(const OrcSim *s, float *obs) {
  int k = 0, task = s->m.task;
  for (int j = 0; j < 6; j++) obs[k++] = (float)s->qpos[j];
  for (int j = 0; j < 6; j++) obs[k++] = (float)s->qvel[j];
  if (task == LCR_TASK_PUSH || task == LCR_TASK_PICK_PLACE) for (int j = 0; j < 3; j++) obs[k++] = (float)s->target[j];
  for (int j = 0; j < 3; j++) obs[k++] = (float)s->qpos[6 + j];
  if (task == LCR_TASK_STACK) for (int j = 0; j < 3; j++) obs[k++] = (float)s->qpos[13 + j];
}

/* Env.reset (reach_cube_env.py:297-311, push_cube_env.py:308-328, lift_cube_env.py:306-320,
 * pick_place_cube_env.py:316-336, stack_two_cubes_env.py:307-324).  qvel/ctrl/warmstart/time are
 * deliberately NOT reset (the reference never calls mj_resetData). */
/* sampling box of PushCubeLoop (push_cube_loop_env.py:133-135): high = geom_size / 2, high[:2] -= 0.008, low = high * (-1, -1, 1) */
static void loop_goal_box(const LcrModel *m, double *lo, double *hi) {
  for (int k = 0; k < 3; k++) hi[k] = m->goal_size[k] / 2;
  hi[0] -= 0.008; hi[1] -= 0.008;
  lo[0] = hi[0] * -1.0; lo[1] = hi[1] * -1.0; lo[2] = hi[2] * 1.0;
}

static void reset_(OrcSim *s, float *obs) {
  const LcrModel *m = &s->m;
  double p[3];
  for (int j = 0; j < 6; j++) s->qpos[j] = 0;
  if (m->task == LCR_TASK_PUSH_LOOP) {
    /* push_cube_loop_env.py:302-320: the cube is drawn inside the region of the CURRENT goal (which persists across
     * resets; kept in target[0]); nothing else is reset */
    double lo[3], hi[3];
    const int cg = s->target[0] != 0;
    loop_goal_box(m, lo, hi);
    draw_uniform3(s->rng, lo, hi, p);
    p[0] += m->goal_center[cg][0]; p[1] += m->goal_center[cg][1]; /* (1 - cg) * c1 + cg * c2 is exact for cg in {0, 1} */
    double *qp = s->qpos + 6;
    qp[0] = p[0]; qp[1] = p[1]; qp[2] = p[2]; qp[3] = 1; qp[4] = qp[5] = qp[6] = 0;
  } else
  for (int c = 0; c < m->ncube; c++) {
    draw_uniform3(s->rng, s->cfg.cube_low, s->cfg.cube_high, p);
    double *qp = s->qpos + 6 + 7 * c;
    qp[0] = p[0]; qp[1] = p[1]; qp[2] = p[2]; qp[3] = 1; qp[4] = qp[5] = qp[6] = 0;
  }
  if (m->task == LCR_TASK_PUSH || m->task == LCR_TASK_PICK_PLACE) {
    draw_uniform3(s->rng, s->cfg.target_low, s->cfg.target_high, p);
    for (int k = 0; k < 3; k++) s->target[k] = (double)(float)p[k]; /* .astype(np.float32), push_cube_env.py:320 */
  }
  forward_(s);
  s->elapsed = 0;
  s->needs_reset = 0;
  if (obs) write_obs(s, obs);
}
void orc_reset(OrcSim *s, float *obs) { reset_(s, obs); }

/* inverse_kinematics + check_joint_limits (reach_cube_env.py:141-221).  Faithful to the reference:
 * each iterate is written to data.qpos and mj_forward is run on it, and the arm is left there. */
static void inverse_kinematics(OrcSim *s, const double *target, double *q_out, int teleport) {
  const LcrModel *m = &s->m;
  double q[6], save[6];
  memcpy(q, s->qpos, sizeof q);
  memcpy(save, s->qpos, sizeof save);
  for (int it = 0; it < 10; it++) {
    memcpy(s->qpos, q, sizeof q);
    if (teleport) forward_(s); else kinematics(s);
    double err[3], en;
    sub3(err, target, s->site_xpos);
    en = norm3(err);
    if (en < 0.01) break;
    double Jc[3][6]; /* mj_jacSite, translational part; the gripper column is zero (site on link_5) */
    for (int j = 0; j < 6; j++) {
      if (j < m->site_body) {
        double r[3], c[3];
        sub3(r, s->site_xpos, s->xpos[j + 1]);
        cross3(c, s->axis[j], r);
        for (int k = 0; k < 3; k++) Jc[k][j] = c[k];
      } else for (int k = 0; k < 3; k++) Jc[k][j] = 0;
    }
    double A[NV][NV], L[NV][NV];
    double rhs[NV], qd[NV];
    for (int a = 0; a < 6; a++) {
      for (int b = 0; b < 6; b++) {
        double t = (a == b) ? 0.15 : 0.0;
        for (int k = 0; k < 3; k++) t += Jc[k][a] * Jc[k][b];
        A[a][b] = t;
      }
      rhs[a] = Jc[0][a] * err[0] + Jc[1][a] * err[1] + Jc[2][a] * err[2];
    }
    cholesky(6, A, L);
    chol_solve(6, L, rhs, qd);
    double n = 0;
    for (int a = 0; a < 6; a++) n += qd[a] * qd[a];
    n = sqrt(n);
    if (n > 1.0) for (int a = 0; a < 6; a++) qd[a] /= n;
    for (int a = 0; a < 6; a++) q[a] = clampd(q[a] + 0.5 * qd[a], m->jnt_range[a][0], m->jnt_range[a][1]);
  }
  memcpy(q_out, q, sizeof q);
  if (!teleport) { memcpy(s->qpos, save, sizeof save); kinematics(s); }
}

void orc_ik(OrcSim *s, const float *target, float *q_out) {
  double t[3] = {target[0], target[1], target[2]}, q[6], site[3], cx[LCR_MAXCUBE][3];
  memcpy(site, s->site_xpos, sizeof site); memcpy(cx, s->cube_xpos, sizeof cx);
  inverse_kinematics(s, t, q, 0);
  memcpy(s->site_xpos, site, sizeof site); memcpy(s->cube_xpos, cx, sizeof cx);
  for (int j = 0; j < 6; j++) q_out[j] = (float)q[j];
}

/* apply_action (reach_cube_env.py:223-279; gripper variants lift_cube_env.py:241-277) */
static void apply_action(OrcSim *s, const float *action_in) {
  const LcrEnvCfg *cfg = &s->cfg;
  const LcrModel *m = &s->m;
  int na = orc_action_dim(cfg), task = m->task;
  int gripper_task = (task == LCR_TASK_LIFT || task == LCR_TASK_PICK_PLACE || task == LCR_TASK_STACK);
  double a[6], tq[6];
  for (int k = 0; k < na; k++) a[k] = clampd((double)action_in[k], -1.0, 1.0);
  if (cfg->action_mode == 1) {
    double tgt[3];
    /* `ee_action * 0.05` is a float32 array times a Python float: the product is rounded to float32 before it is added to
     * the float64 site position (reach_cube_env.py:241; NumPy promotion) */
    for (int k = 0; k < 3; k++) tgt[k] = s->site_xpos[k] + (double)((float)a[k] * 0.05f);
    if (tgt[2] < 0) tgt[2] = 0;
    inverse_kinematics(s, tgt, tq, 1);
    if (!gripper_task) tq[5] = 0;
    else {
      /* action[3] (lift_cube_env.py:242); with block_gripper=True the action has 3 entries and the
       * reference would raise IndexError -- we use 0 for the missing gripper entry */
      double ga = na > 3 ? a[3] : 0.0;
      /* `gripper_action * 0.2`: np.float32 scalar times a Python float stays float32 (NumPy >= 2, NEP 50; lift_cube_env.py:253) */
      tq[5] = clampd(s->qpos[5] + (double)((float)ga * 0.2f), m->act_ctrlrange[5][0], m->act_ctrlrange[5][1]);
    }
  } else {
    for (int j = 0; j < 5; j++) tq[j] = clampd(a[j] + s->qpos[j], TARGET_LOW[j], TARGET_HIGH[j]);
    if (!gripper_task) tq[5] = 0;
    else tq[5] = clampd(a[na - 1] + s->qpos[5], TARGET_LOW[5], TARGET_HIGH[5]); /* action[-1], lift_cube_env.py:264 */
  }
  memcpy(s->ctrl, tq, sizeof tq);
  for (int k = 0; k < cfg->n_substeps; k++) substep_(s);
}

/* Env.step (reach_cube_env.py:313-348, push_cube_env.py:330-361, lift_cube_env.py:322-346,
 * pick_place_cube_env.py:338-369, stack_two_cubes_env.py:326-363) + TimeLimit */
/* get_reward / get_cube_overlap of PushCubeLoop (push_cube_loop_env.py:337-383).  The reference mixes numpy float32
 * (cube_position), numpy float64 (model arrays) and Python scalars; the dtype of every intermediate follows NumPy >= 2
 * promotion (NEP 50: a Python scalar adopts the dtype of the numpy operand), tracked explicitly here. */
typedef struct { double v; int k; } PyNum; /* k: 0 numpy float64, 1 numpy float32, 2 Python scalar (weak) */
static PyNum pn(double v, int k) { PyNum r = {v, k}; return r; }
static PyNum pn_op(PyNum a, PyNum b, int op) { /* 0 +, 1 -, 2 *, 3 / */
  const int k = (a.k == 0 || b.k == 0) ? 0 : ((a.k == 1 || b.k == 1) ? 1 : 2);
  if (k == 1) {
    const float x = (float)a.v, y = (float)b.v;
    const float z = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : x / y;
    return pn((double)z, 1);
  }
  return pn(op == 0 ? a.v + b.v : op == 1 ? a.v - b.v : op == 2 ? a.v * b.v : a.v / b.v, k);
}
static PyNum pn_min(PyNum a, PyNum b) { return b.v < a.v ? b : a; } /* Python min / max return the first extremal operand */
static PyNum pn_max(PyNum a, PyNum b) { return b.v > a.v ? b : a; }

static void loop_reward(OrcSim *s, float *reward, uint8_t *success) {
  const LcrModel *m = &s->m;
  const int cg = s->target[0] != 0;
  double lo[3], hi[3];
  loop_goal_box(m, lo, hi);
  const PyNum w = pn(0.015 / 2, 2); /* self.cube_size (push_cube_loop_env.py:124) */
  PyNum ov[2];
  for (int k = 0; k < 2; k++) {
    const PyNum c = pn((double)(float)s->qpos[6 + k], 1); /* data.qpos[6:9].astype(np.float32) */
    const PyNum g = pn(m->goal_center[cg][k], 0), wg = pn(hi[k], 0);
    const PyNum up = pn_min(pn_op(c, w, 0), pn_op(g, wg, 0)), dn = pn_max(pn_op(c, w, 1), pn_op(g, wg, 1));
    ov[k] = pn_max(pn(0, 2), pn_op(up, dn, 1));
  }
  const PyNum area = pn_op(ov[0], ov[1], 2);
  const PyNum cube_area = pn(0.0075 * 0.0075 * 4, 2); /* w_cube * l_cube * 4, Python floats */
  const PyNum overlap = pn_op(area, cube_area, 3);
  *success = 0;
  if (overlap.v > 0.95) { /* success: +5 and the goal switches */
    *success = 1;
    *reward = 5.0f;
    s->target[0] = cg ? 0.0 : 1.0;
  } else if (overlap.v > 0.0) {
    *reward = (float)pn_op(overlap, pn(1, 2), 1).v;
  } else { /* distance to the near edge of the goal region along y only */
    const double edge = lo[1] + m->goal_center[cg][1];
    const double diff = (double)(float)s->qpos[7] - edge, dist = sqrt(diff * diff);
    double r = (-dist / 0.16) - 1;
    r = r > -2 ? r : -2; /* max(r, -2) */
    r = -1 < r ? -1 : r; /* min(., -1) */
    *reward = (float)r;
  }
}

void orc_step(OrcSim *s, const float *action, float *obs, float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *success) {
  const LcrEnvCfg *cfg = &s->cfg;
  int task = s->m.task;
  if (cfg->autoreset && s->needs_reset) {
    reset_(s, obs);
    *reward = 0; *terminated = 0; *truncated = 0; *success = 0;
    return;
  }
  s->max_nefc = 0; s->step_max_ncon = 0;
  apply_action(s, action);
  { int b = s->max_nefc / 16; s->hist_nefc[b < 31 ? b : 31]++; b = s->step_max_ncon / 8; s->hist_ncon[b < 31 ? b : 31]++; }
  write_obs(s, obs);
  if (task == LCR_TASK_PUSH_LOOP) { /* push_cube_loop_env.py:322-335: never terminates; TimeLimit truncates */
    loop_reward(s, reward, success);
    *terminated = 0;
    s->elapsed++;
    *truncated = (cfg->max_episode_steps > 0 && s->elapsed >= cfg->max_episode_steps);
    s->needs_reset = *truncated;
    return;
  }
  double d = 0, a[3], b[3];
  if (task == LCR_TASK_REACH || task == LCR_TASK_LIFT) { memcpy(a, s->site_xpos, sizeof a); memcpy(b, s->cube_xpos[0], sizeof b); }
  else if (task == LCR_TASK_STACK) { memcpy(a, s->cube_xpos[1], sizeof a); memcpy(b, s->cube_xpos[0], sizeof b); b[2] += 0.03; }
  else { memcpy(a, s->cube_xpos[0], sizeof a); memcpy(b, s->target, sizeof b); }
  sub3(a, a, b);
  d = norm3(a);
  if (task == LCR_TASK_LIFT) {
    *reward = (float)((s->cube_xpos[0][2] - cfg->height_threshold) + d);
    *terminated = 0; *success = 0;
  } else {
    *success = d < cfg->distance_threshold;
    *terminated = *success;
    *reward = cfg->reward_type == 0 ? -(float)(d > cfg->distance_threshold) : (float)(-d);
  }
  s->elapsed++;
  *truncated = (cfg->max_episode_steps > 0 && s->elapsed >= cfg->max_episode_steps);
  s->needs_reset = (*terminated || *truncated);
}

/* ------------------------------------------------------------------ handle management / accessors */
OrcSim *orc_create(const LcrModel *m, const double *verts, const LcrEnvCfg *cfg) {
  OrcSim *s = (OrcSim *)calloc(1, sizeof(OrcSim));
  s->m = *m; s->cfg = *cfg;
  s->sw_tilt = 1e-3; s->sw_kv_clamped = 1; s->sw_impratio = 0;
  s->verts = (double *)malloc(sizeof(double) * 3 * m->nvert);
  memcpy(s->verts, verts, sizeof(double) * 3 * m->nvert);
  reset_data(s);
  sa_clear(s);
  s->rng[1] = 1; s->rng[3] = 1;
  return s;
}
void orc_destroy(OrcSim *s) { if (s) { free(s->verts); free(s); } }
void orc_seed(OrcSim *s, const uint64_t *st) { memcpy(s->rng, st, sizeof s->rng); }
void orc_get_rng(OrcSim *s, uint64_t *st) { memcpy(st, s->rng, sizeof s->rng); }
double orc_rng_double(OrcSim *s) { return pcg64_double(s->rng); }

void orc_get_state(const OrcSim *s, double *qpos, double *qvel, double *ctrl, double *warm, double *aux, int32_t *ints) {
  if (qpos) memcpy(qpos, s->qpos, sizeof(double) * s->m.nq);
  if (qvel) memcpy(qvel, s->qvel, sizeof(double) * s->m.nv);
  if (ctrl) memcpy(ctrl, s->ctrl, sizeof(double) * 6);
  if (warm) memcpy(warm, s->warm, sizeof(double) * s->m.nv);
  if (aux) {
    aux[0] = s->time; memcpy(aux + 1, s->target, 24); memcpy(aux + 4, s->site_xpos, 24);
    memcpy(aux + 7, s->cube_xpos, 48);
  }
  if (ints) { ints[0] = s->elapsed; ints[1] = s->needs_reset; }
}
void orc_set_state(OrcSim *s, const double *qpos, const double *qvel, const double *ctrl, const double *warm, const double *aux, const int32_t *ints) {
  if (qpos) memcpy(s->qpos, qpos, sizeof(double) * s->m.nq);
  if (qvel) memcpy(s->qvel, qvel, sizeof(double) * s->m.nv);
  if (ctrl) memcpy(s->ctrl, ctrl, sizeof(double) * 6);
  if (warm) memcpy(s->warm, warm, sizeof(double) * s->m.nv);
  if (aux) {
    s->time = aux[0]; memcpy(s->target, aux + 1, 24); memcpy(s->site_xpos, aux + 4, 24);
    memcpy(s->cube_xpos, aux + 7, 48);
  }
  if (ints) { s->elapsed = ints[0]; s->needs_reset = ints[1]; }
  sa_clear(s); /* the separating-axis cache is not part of the checkpointed state */
}
/* lifetime peaks: contacts, constraint rows, convex candidates of one substep; total dropped contacts */
int orc_set_switch(OrcSim *s, const char *name, double value) {
  if (!strcmp(name, "plane_hull_tilt")) s->sw_tilt = value;
  else if (!strcmp(name, "implicit_kv_when_clamped")) s->sw_kv_clamped = value != 0;
  else if (!strcmp(name, "impratio")) s->sw_impratio = value;
  else return -1;
  return 0;
}
void orc_get_peaks(const OrcSim *s, int32_t *d) { d[0] = s->peak_ncon; d[1] = s->peak_nefc; d[2] = s->peak_ncand; d[3] = s->total_overflow; }
void orc_get_hist(const OrcSim *s, int32_t *nefc16, int32_t *ncon8) { memcpy(nefc16, s->hist_nefc, sizeof s->hist_nefc); memcpy(ncon8, s->hist_ncon, sizeof s->hist_ncon); }
void orc_get_diag(const OrcSim *s, int32_t *d) {
  d[0] = s->ncon; d[1] = s->nefc; d[2] = s->niter; d[3] = s->max_nefc; d[4] = s->overflow; d[5] = s->nan_resets;
}

/* named read-only views of the last forward pass, for the known-answer tests */
int orc_get(const OrcSim *s, const char *name, double *out, int cap) {
  const double *src = NULL;
  int n = 0, nv = s->m.nv;
  static double tmp[LCR_MAXEFC_BIG * NV > LCR_MAXCON_BIG * 32 ? LCR_MAXEFC_BIG * NV : LCR_MAXCON_BIG * 32];
  const int nbd = LCR_NABODY + LCR_MAXCUBE; /* the dynamic bodies (the wall slots behind them are constants) */
  if (!strcmp(name, "xpos")) { src = &s->xpos[0][0]; n = nbd * 3; }
  else if (!strcmp(name, "xmat")) { src = &s->xmat[0][0]; n = nbd * 9; }
  else if (!strcmp(name, "xipos")) { src = &s->xipos[0][0]; n = nbd * 3; }
  else if (!strcmp(name, "axis")) { src = &s->axis[0][0]; n = 18; }
  else if (!strcmp(name, "site_xpos")) { src = s->site_xpos; n = 3; }
  else if (!strcmp(name, "M")) { for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) tmp[i * nv + j] = s->M[i][j]; src = tmp; n = nv * nv; }
  else if (!strcmp(name, "bias")) { src = s->bias; n = nv; }
  else if (!strcmp(name, "passive")) { src = s->passive; n = nv; }
  else if (!strcmp(name, "actuator")) { src = s->actuator; n = nv; }
  else if (!strcmp(name, "qacc_smooth")) { src = s->qacc_smooth; n = nv; }
  else if (!strcmp(name, "qacc")) { src = s->qacc; n = nv; }
  else if (!strcmp(name, "qfrc_constraint")) { src = s->qfrc_constraint; n = nv; }
  else if (!strcmp(name, "efc_J")) { for (int i = 0; i < s->nefc; i++) for (int j = 0; j < nv; j++) tmp[i * nv + j] = s->J[i][j]; src = tmp; n = s->nefc * nv; }
  else if (!strcmp(name, "efc_pos")) { src = s->efc_pos; n = s->nefc; }
  else if (!strcmp(name, "efc_R")) { src = s->efc_R; n = s->nefc; }
  else if (!strcmp(name, "efc_D")) { src = s->efc_D; n = s->nefc; }
  else if (!strcmp(name, "efc_aref")) { src = s->efc_aref; n = s->nefc; }
  else if (!strcmp(name, "efc_force")) { src = s->efc_force; n = s->nefc; }
  else if (!strcmp(name, "efc_vel")) { src = s->efc_vel; n = s->nefc; }
  else if (!strcmp(name, "contacts")) { /* per contact: pos3 frame9 dist dim g1 g2 mu friction5 solimp5 = 27 */
    for (int i = 0; i < s->ncon; i++) {
      const Contact *c = &s->con[i];
      double *t = tmp + 27 * i;
      memcpy(t, c->pos, 24); memcpy(t + 3, c->frame, 72); t[12] = c->dist; t[13] = c->dim; t[14] = c->g1; t[15] = c->g2;
      t[16] = c->mu; memcpy(t + 17, c->friction, 40); memcpy(t + 22, c->solimp, 40);
    }
    src = tmp; n = 27 * s->ncon;
  } else return -1;
  if (n > cap) n = cap;
  memcpy(out, src, sizeof(double) * n);
  return n;
}

/* run `nsteps` env steps on `nenv` independent sims with a shared action tensor [nsteps][nenv][na];
 * used by bench.py's cpu_baseline leg (one call per worker thread, disjoint sims) */
void orc_rollout(OrcSim **sims, int nenv, int nsteps, const float *actions, float *obs_last, float *reward_sum) {
  int na = orc_action_dim(&sims[0]->cfg), od = orc_obs_dim(sims[0]->m.task);
  float obs[18], r; uint8_t te, tr, su;
  for (int e = 0; e < nenv; e++) {
    double acc = 0;
    for (int t = 0; t < nsteps; t++) {
      orc_step(sims[e], actions + ((size_t)t * nenv + e) * na, obs, &r, &te, &tr, &su);
      acc += r;
    }
    if (obs_last) memcpy(obs_last + (size_t)e * od, obs, sizeof(float) * od);
    if (reward_sum) reward_sum[e] = (float)acc;
  }
}

int orc_sizeof_model(void) { return (int)sizeof(LcrModel); }
int orc_sizeof_cfg(void) { return (int)sizeof(LcrEnvCfg); }
