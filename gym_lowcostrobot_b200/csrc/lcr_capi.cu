// lcr_capi.cu -- host side of liblcrsim.so: the C-ABI declared in include/lcrsim.h.
// Builds the device model (incl. the contact-parameter classes from MuJoCo's geom mixing rule),
// owns the SoA state buffers, and enqueues the kernels of lcr_kernels.cuh.  No CPU fallback: if
// CUDA is unavailable every call fails with an error.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lcr_device.cuh"

namespace {

thread_local std::string g_err;
int fail(const std::string& msg) { g_err = msg; return 1; }
#define CUDA_OK(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

constexpr double kMinVal = 1e-15, kMinImp = 1e-4, kMaxImp = 0.9999, kMinMu = 1e-5;

struct Mixed {
  int dim;
  double fr[3], solref[2], solimp[5];
};

// MuJoCo contact parameter mixing: higher priority wins; on a tie max condim, element-wise max
// friction, solmix-weighted solref / solimp.
Mixed mix(const LcrModel& m, int g1, int g2) {
  Mixed o;
  const int p1 = m.geom_priority[g1], p2 = m.geom_priority[g2];
  if (p1 != p2) {
    const int g = p1 > p2 ? g1 : g2;
    o.dim = m.geom_condim[g];
    std::copy(m.geom_friction[g], m.geom_friction[g] + 3, o.fr);
    std::copy(m.geom_solref[g], m.geom_solref[g] + 2, o.solref);
    std::copy(m.geom_solimp[g], m.geom_solimp[g] + 5, o.solimp);
    return o;
  }
  o.dim = std::max(m.geom_condim[g1], m.geom_condim[g2]);
  for (int k = 0; k < 3; k++) o.fr[k] = std::max(m.geom_friction[g1][k], m.geom_friction[g2][k]);
  const double s1 = m.geom_solmix[g1], s2 = m.geom_solmix[g2];
  double w;
  if (s1 >= kMinVal && s2 >= kMinVal) w = s1 / (s1 + s2);
  else if (s1 < kMinVal && s2 < kMinVal) w = 0.5;
  else w = s1 < kMinVal ? 0.0 : 1.0;
  const bool pos = m.geom_solref[g1][0] > 0 && m.geom_solref[g2][0] > 0;
  for (int k = 0; k < 2; k++)
    o.solref[k] = pos ? w * m.geom_solref[g1][k] + (1 - w) * m.geom_solref[g2][k] : std::min(m.geom_solref[g1][k], m.geom_solref[g2][k]);
  for (int k = 0; k < 5; k++) o.solimp[k] = w * m.geom_solimp[g1][k] + (1 - w) * m.geom_solimp[g2][k];
  return o;
}

template <typename T>
void fill_par(CPar<T>& p, int dim, const double* fr3, const double* solref, const double* solimp, double timestep) {
  const double f[5] = {std::max(kMinMu, fr3[0]), std::max(kMinMu, fr3[0]), std::max(kMinMu, fr3[1]), std::max(kMinMu, fr3[2]),
                       std::max(kMinMu, fr3[2])};
  for (int k = 0; k < 5; k++) p.fr[k] = (T)f[k];
  const double si[5] = {std::min(std::max(solimp[0], kMinImp), kMaxImp), std::min(std::max(solimp[1], kMinImp), kMaxImp),
                        std::max(0.0, solimp[2]), std::min(std::max(solimp[3], kMinImp), kMaxImp), std::max(1.0, solimp[4])};
  for (int k = 0; k < 5; k++) p.si[k] = (T)si[k];
  const double dmax = si[1];
  double K, B;
  if (solref[0] > 0) {
    const double tc = std::max(solref[0], 2 * timestep), dr = solref[1];  // refsafe
    K = 1 / std::max(kMinVal, dmax * dmax * tc * tc * dr * dr);
    B = 2 / std::max(kMinVal, dmax * tc);
  } else {
    K = -solref[0] / std::max(kMinVal, dmax * dmax);
    B = -solref[1] / std::max(kMinVal, dmax);
  }
  p.K = (T)K; p.B = (T)B; p.dim = dim;
}

template <typename T>
void build_model(const LcrModel& m, const LcrEnvCfg& c, DevModel<T>& d) {
  std::memset(&d, 0, sizeof d);
  d.task = m.task; d.ncube = m.ncube; d.nq = m.nq; d.nv = m.nv; d.nmesh = m.nmesh; d.nvert = m.nvert; d.npair = m.npair;
  d.site_body = m.site_body; d.iterations = m.iterations; d.ls_iterations = m.ls_iterations;
  d.timestep = (T)m.timestep; d.impratio = (T)m.impratio; d.tolerance = (T)m.tolerance; d.ls_tolerance = (T)m.ls_tolerance;
  d.meaninertia = (T)m.meaninertia;
  for (int k = 0; k < 3; k++) { d.gravity[k] = (T)m.gravity[k]; d.site_pos[k] = (T)m.site_pos[k]; }
  for (int b = 0; b < LCR_NABODY; b++) {
    for (int k = 0; k < 3; k++) { d.body_pos[b][k] = (T)m.body_pos[b][k]; d.body_ipos[b][k] = (T)m.body_ipos[b][k]; d.body_inertia[b][k] = (T)m.body_inertia[b][k]; }
    for (int k = 0; k < 4; k++) { d.body_quat[b][k] = (T)m.body_quat[b][k]; d.body_iquat[b][k] = (T)m.body_iquat[b][k]; }
    d.body_mass[b] = (T)m.body_mass[b];
    d.body_invweight0[b][0] = (T)m.body_invweight0[b][0]; d.body_invweight0[b][1] = (T)m.body_invweight0[b][1];
  }
  for (int cb = 0; cb < m.ncube; cb++) {
    d.body_invweight0[LCR_NABODY + cb][0] = (T)m.cube_invweight0[cb][0]; d.body_invweight0[LCR_NABODY + cb][1] = (T)m.cube_invweight0[cb][1];
    d.cube_mass[cb] = (T)m.cube_mass[cb]; d.cube_inertia[cb] = (T)m.cube_inertia[cb][0];
    for (int k = 0; k < 3; k++) { d.cube_size[cb][k] = (T)m.cube_size[cb][k]; d.cube_qpos0[cb][k] = (T)m.cube_pos0[cb][k]; }
  }
  for (int j = 0; j < LCR_NARM; j++) {
    for (int k = 0; k < 3; k++) d.jnt_axis[j][k] = (T)m.jnt_axis[j][k];
    for (int k = 0; k < 2; k++) { d.jnt_range[j][k] = (T)m.jnt_range[j][k]; d.jnt_frcrange[j][k] = (T)m.jnt_frcrange[j][k]; d.act_ctrlrange[j][k] = (T)m.act_ctrlrange[j][k]; }
    d.jnt_armature[j] = (T)m.jnt_armature[j]; d.jnt_damping[j] = (T)m.jnt_damping[j]; d.dof_invweight0[j] = (T)m.dof_invweight0[j];
    d.act_kp[j] = (T)m.act_kp[j]; d.act_kv[j] = (T)m.act_kv[j];
    const double fr[3] = {1, 0.005, 0.0001};
    fill_par(d.par_limit[j], 1, fr, m.jnt_solref[j], m.jnt_solimp[j], m.timestep);
  }
  for (int g = 0; g < m.nmesh; g++) {
    d.mesh_body[g] = m.mesh_body[g]; d.mesh_vertadr[g] = m.mesh_vertadr[g]; d.mesh_vertnum[g] = m.mesh_vertnum[g];
    for (int k = 0; k < 3; k++) { d.mesh_center[g][k] = (T)m.mesh_center[g][k]; d.mesh_half[g][k] = (T)m.mesh_half[g][k]; d.mesh_com[g][k] = (T)m.mesh_com[g][k]; }
    d.mesh_rbound[g] = (T)m.mesh_rbound[g];
  }
  const int gfloor = m.nmesh;
  for (int cb = 0; cb < m.ncube; cb++) {
    Mixed x = mix(m, gfloor, gfloor + 1 + cb);
    fill_par(d.par_floor_cube[cb], x.dim, x.fr, x.solref, x.solimp, m.timestep);
    for (int g = 0; g < m.nmesh; g++) {
      // geom order follows MuJoCo's type ordering: box (cube) is geom1, mesh is geom2
      Mixed y = mix(m, gfloor + 1 + cb, g);
      fill_par(d.par_cube_mesh[cb][g], y.dim, y.fr, y.solref, y.solimp, m.timestep);
    }
  }
  if (m.ncube == 2) {
    Mixed x = mix(m, gfloor + 1, gfloor + 2);
    fill_par(d.par_cube_cube, x.dim, x.fr, x.solref, x.solimp, m.timestep);
  }
  // static wall boxes (PushCubeLoop): box index ncube + w, geom gfloor + 1 + ncube + w
  for (int wi = 0; wi < m.nwall; wi++) {
    const int c = m.ncube + wi, gw = gfloor + 1 + c;
    for (int k = 0; k < 3; k++) { d.cube_size[c][k] = (T)m.wall_size[wi][k]; d.wall_pos[wi][k] = (T)m.wall_pos[wi][k]; }
    for (int cb = 0; cb < m.ncube; cb++) {
      Mixed x = mix(m, gw, gfloor + 1 + cb);
      fill_par(d.par_wall_cube[wi][cb], x.dim, x.fr, x.solref, x.solimp, m.timestep);
    }
    for (int g = 0; g < m.nmesh; g++) {
      Mixed y = mix(m, gw, g);
      fill_par(d.par_cube_mesh[c][g], y.dim, y.fr, y.solref, y.solimp, m.timestep);
    }
  }
  for (int k = 0; k < 3; k++) {  // push_cube_loop_env.py:127-135, same float64 operations as numpy
    d.goal_center[0][k] = m.goal_center[0][k]; d.goal_center[1][k] = m.goal_center[1][k];
    d.goal_high[k] = m.goal_size[k] / 2;
    if (k < 2) d.goal_high[k] -= 0.008;
  }
  for (int g = 0; g < m.nmesh; g++) {
    Mixed x = mix(m, gfloor, g);
    fill_par(d.par_floor_mesh[g], x.dim, x.fr, x.solref, x.solimp, m.timestep);
  }
  for (int p = 0; p < m.npair; p++) {
    d.pair_g1[p] = m.pair_g1[p]; d.pair_g2[p] = m.pair_g2[p];
    Mixed x = mix(m, m.pair_g1[p], m.pair_g2[p]);
    fill_par(d.par_mesh_mesh[p], x.dim, x.fr, x.solref, x.solimp, m.timestep);
  }
  d.action_mode = c.action_mode; d.block_gripper = c.block_gripper; d.reward_type = c.reward_type; d.n_substeps = c.n_substeps;
  d.max_episode_steps = c.max_episode_steps; d.autoreset = c.autoreset; d.collision_mask = c.collision_mask;
  d.distance_threshold = (T)c.distance_threshold; d.height_threshold = (T)c.height_threshold;
  for (int k = 0; k < 3; k++) { d.cube_low[k] = c.cube_low[k]; d.cube_high[k] = c.cube_high[k]; d.target_low[k] = c.target_low[k]; d.target_high[k] = c.target_high[k]; }
}

// dispatch on the scene class (S = 1, 2 cubes or LCR_NC_LOOP)
#define LCR_DISPATCH(T, nc, CALL)                          \
  do {                                                     \
    if ((nc) == 1) lcr::LaunchNC<T, 1>::CALL;              \
    else if ((nc) == 2) lcr::LaunchNC<T, 2>::CALL;         \
    else lcr::LaunchNC<T, LCR_NC_LOOP>::CALL;              \
  } while (0)
#define LCR_DISPATCH_RET(T, nc, out, CALL)                 \
  do {                                                     \
    if ((nc) == 1) out = lcr::LaunchNC<T, 1>::CALL;        \
    else if ((nc) == 2) out = lcr::LaunchNC<T, 2>::CALL;   \
    else out = lcr::LaunchNC<T, LCR_NC_LOOP>::CALL;        \
  } while (0)

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T>
struct Impl {
  DevModel<T>* dm = nullptr;
  T* verts = nullptr;
  DevState<T> s{};
  void* gws = nullptr;  // per-env parked workspaces of the phased and flow execution modes
  int nf = 0, nc = 0;

  int create(const LcrModel& m, const double* hv, const LcrEnvCfg& c, int n) {
    DevModel<T> h;
    build_model(m, c, h);
    CUDA_OK(cudaMalloc(&dm, sizeof h));
    CUDA_OK(cudaMemcpy(dm, &h, sizeof h, cudaMemcpyHostToDevice));
    std::vector<T> v(4 * (size_t)std::max(m.nvert, 1), (T)0);
    for (int i = 0; i < m.nvert; i++) for (int k = 0; k < 3; k++) v[4 * (size_t)i + k] = (T)hv[3 * (size_t)i + k];
    CUDA_OK(cudaMalloc(&verts, v.size() * sizeof(T)));
    CUDA_OK(cudaMemcpy(verts, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    nf = m.nq + 2 * m.nv + LCR_NARM + LCR_NAUX;
    s.n = n;
    s.nfp = (nf + 3) & ~3;
    CUDA_OK(cudaMalloc(&s.st, sizeof(T) * (size_t)s.nfp * n));
    CUDA_OK(cudaMalloc(&s.ib, sizeof(int32_t) * LCR_IB_WORDS * (size_t)n));
    CUDA_OK(cudaMalloc(&s.sa, (size_t)Ws<T, 1>::SA_BYTES * n));
    nc = scene_class(m.task, m.ncube);
    if (c.exec_mode == 1 || c.exec_mode == 3) {
      size_t wsb = 0;
      LCR_DISPATCH_RET(T, nc, wsb, ws_bytes());
      CUDA_OK(cudaMalloc(&gws, wsb * (size_t)n));
    }
    LCR_DISPATCH(T, nc, prepare());
    lcr::Launch<T>::init_state(dm, s, 0);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return 0;
  }
  void destroy() {
    cudaFree(dm); cudaFree(verts); cudaFree(s.st); cudaFree(s.ib); cudaFree(s.sa); cudaFree(gws);
  }
};

}  // namespace

struct LcrSim {
  int precision = 0, device = 0, n = 0, ncube = 0, task = 0, launches = 0;
  // phased mode: the env range is cut into groups, each with its own stream, so that the tail of one group's
  // variable-cost kernels (collision, solver) overlaps the other groups' work
  int ngroups = 0;
  long long* prof = nullptr;  // debug: per-env phase clocks of the last lockstep step, see lcr_debug_phase_clocks
  int* perm = nullptr;  // lockstep / phased mode: work-aware env order of the current step (device)
  int ls_striped = 1;  // seat order of the lockstep scheduler, see k_sched
  int ls_warps = 0, ls_flags = 0;  // lockstep mode: envs per CTA (0 = as many as fit one SM) and LCR_LS_* barrier flags
  cudaStream_t gstream[16] = {};
  cudaEvent_t ev_begin = nullptr, ev_done[16] = {};
  // envs that outgrew the fast workspace in the current call: device counter + list, redone over the big workspace
  int* redo = nullptr;
  // envs that start over the big workspace (predicted by the scheduler from their previous step): count + list, own stream
  int* big = nullptr;
  int tbig = 0;
  cudaStream_t bstream = nullptr;
  cudaEvent_t ev_sched = nullptr, ev_big = nullptr;
  // flow mode: queue control block + rings, grid, tunables
  FlowQ fq{};
  unsigned* fq_mem = nullptr;
  unsigned long long* fq_rings = nullptr;
  int flow_grid = 0, flow_bigcta = 0, flow_flags = 0, flow_thi = 0, flow_tbig = 0;
  unsigned long long* flow_stats = nullptr;  // debug: busy clocks per phase of the flow kernel, see lcr_debug_flow_stats
  unsigned long long* seed_buf = nullptr;    // device staging of lcr_seed
  // phased mode: the whole step (scheduler, BIG branch, the chains of all env groups, redo pass) is captured once into a CUDA
  // graph and replayed with one launch per step; the actions are staged into act_buf so that the captured kernel arguments
  // do not depend on the caller's action pointer; the graph is re-captured if the caller's output pointers change
  int use_graph = 0, ph_striped = 1, graph_nodes = 0;
  // phased mode: migration lists [group][substep][1 + LCR_MIGCAP] and the side stream / fork / join events of every (group, substep)
  int* mig = nullptr;
  int* early = nullptr;  // phased mode: per group [1 + per] list of the envs whose step begin outgrew the fast workspace
  int early_stride = 0;
  int* jobq = nullptr;  // phased mode: narrowphase job queues, one per group (jobq_stride ints each)
  int jobq_stride = 0;
  std::vector<cudaStream_t> mstream;
  std::vector<cudaEvent_t> mev_fork, mev_join;
  cudaStream_t cstream = nullptr;
  cudaGraphExec_t gexec = nullptr;
  const void* gkey[6] = {};
  float* act_buf = nullptr;
  LcrModel model;
  LcrEnvCfg cfg;
  Impl<float> f;
  Impl<double> d;
};

#define WITH_DEVICE(sim)                                  \
  if (!(sim)) return fail("null handle");                 \
  CUDA_OK(cudaSetDevice((sim)->device));

// run CALL (a LaunchNC member call) in the sim's precision and scene class
#define LCR_RUN(sim, CALL_F, CALL_D)                                          \
  do {                                                                        \
    if ((sim)->precision == LCR_F32) LCR_DISPATCH(float, (sim)->ncube, CALL_F); \
    else LCR_DISPATCH(double, (sim)->ncube, CALL_D);                          \
  } while (0)

namespace {
Redo redo_of(const LcrSim* s) { return Redo{s->redo, s->redo + 1}; }

unsigned next_pow2(size_t x) { unsigned p = 1; while (p < x) p <<= 1; return p; }

int create_flow(LcrSim* s) {
  if (s->n >= LCR_FQ_MAXENV) return fail("lcr_create: flow mode holds at most 2^20 - 1 envs per device");
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, s->device));
  s->flow_grid = std::max(2, env_int("LCR_FLOW_GRID", prop.multiProcessorCount));
  // BIG CTAs: the envs that need the big workspace are ~1e-3 of the batch but each keeps a warp busy for the whole step
  // (measured on B200, PushCube 16 384 mid-episode: 4 BIG CTAs 41.9 ms / step, 8: 38.4, 16: 36.8 -- the envs on the BIG path are the
  //  tail of the step, each keeps one warp busy for 10 - 20 ms)
  s->flow_bigcta = std::max(0, std::min(s->flow_grid - 1, env_int("LCR_FLOW_BIGCTA", s->n >= 12288 ? 12 : (s->n >= 2048 ? 6 : 2))));  // (0: debug only)
  s->flow_flags = env_int("LCR_FLOW_FLAGS", 1);
  if (env_int("LCR_STACK", 0) > 0) CUDA_OK(cudaDeviceSetLimit(cudaLimitStackSize, (size_t)env_int("LCR_STACK", 0)));  // debug
  s->flow_thi = env_int("LCR_FLOW_THI", 40);
  // envs start on the BIG path only if their previous step ended far beyond the fast caps (measured: predicting at 88 rows sent
  // envs there that would not have overflowed again: 41.9 ms / step against 39.2 with migration alone)
  s->flow_tbig = env_int("LCR_FLOW_TBIG", 112);
  // rings: one per phase and priority, n entries each; the JOB rings hold one item per convex candidate
  // (64-bit slots seq << 32 | item; slot i starts free for ticket i; twice the items that can be outstanding, so that a
  // producer practically never waits for the consumer of the previous lap)
  size_t slots = 0, off[LCR_FQ_NQ];
  for (int q = 0; q < LCR_FQ_NQ; q++) {
    const size_t cap = next_pow2(2 * ((q / 2 == FQ_JOB) ? (size_t)s->n * (3 * LCR_MAXEFC / 8) : (size_t)s->n));
    s->fq.mask[q] = (unsigned)(cap - 1);
    off[q] = slots;
    slots += cap;
  }
  CUDA_OK(cudaMalloc(&s->fq_mem, LCR_FQ_CTL_WORDS * sizeof(unsigned)));
  CUDA_OK(cudaMemset(s->fq_mem, 0, LCR_FQ_CTL_WORDS * sizeof(unsigned)));
  CUDA_OK(cudaMalloc(&s->fq_rings, slots * sizeof(unsigned long long)));
  {
    std::vector<unsigned long long> init(slots);
    for (int q = 0; q < LCR_FQ_NQ; q++)
      for (size_t i = 0; i <= s->fq.mask[q]; i++) init[off[q] + i] = (unsigned long long)i << 32;
    CUDA_OK(cudaMemcpy(s->fq_rings, init.data(), slots * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  }
  s->fq.ctl = s->fq_mem;
  for (int q = 0; q < LCR_FQ_NQ; q++) s->fq.ring[q] = reinterpret_cast<unsigned*>(s->fq_rings + off[q]);
  return 0;
}
}  // namespace

namespace {
// the kernels of one control step in the fused / phased / lockstep modes, enqueued on `st` (and on the sim's side streams, which
// fork from and join `st` through events -- so the same code is what a stream capture records)
int enqueue_step(LcrSim* sim, const StepIO& io, cudaStream_t st) {
  const Redo redo = redo_of(sim);
  const bool f32 = sim->precision == LCR_F32;
  CUDA_OK(cudaMemsetAsync(sim->redo, 0, sizeof(int), st));
  bool predicted = false, big_started = false;
  // the envs the scheduler sent to the big workspace start now, on their own stream, beside the main kernels
  auto start_big = [&]() {
    cudaEventRecord(sim->ev_sched, st);
    cudaStreamWaitEvent(sim->bstream, sim->ev_sched, 0);
    const Redo lst{sim->big, sim->big + 1};
    LCR_RUN(sim, step_big(sim->f.dm, sim->f.verts, sim->f.s, io, lst, sim->bstream), step_big(sim->d.dm, sim->d.verts, sim->d.s, io, lst, sim->bstream));
    cudaEventRecord(sim->ev_big, sim->bstream);
    sim->launches++;
    big_started = true;
  };
  if (sim->cfg.exec_mode == 1) {
    const int G = sim->ngroups, per = (sim->n + G - 1) / G;
    if (sim->perm) {  // seats in work-aware order: group g owns perm[g * per, (g + 1) * per), heaviest envs first, -1 = padding
      if (f32) lcr::Launch<float>::sched(sim->f.s, sim->perm, per, sim->ph_striped, sim->big, sim->tbig, st);
      else lcr::Launch<double>::sched(sim->d.s, sim->perm, per, sim->ph_striped, sim->big, sim->tbig, st);
      sim->launches++;
      predicted = true;
    }
    if (predicted) start_big();
    const int nsub = std::max(1, sim->cfg.n_substeps);
    if (sim->mig) {
      CUDA_OK(cudaMemsetAsync(sim->mig, 0, sizeof(int) * (size_t)G * (nsub + 1) * (1 + LCR_MIGCAP), st));
      CUDA_OK(cudaMemsetAsync(sim->early, 0, sizeof(int) * (size_t)sim->early_stride * G, st));
    }
    CUDA_OK(cudaEventRecord(sim->ev_begin, st));
    for (int g = 0; g < G; g++) {
      const int env0 = g * per, cnt = sim->perm ? per : std::min(per, sim->n - env0);
      if (env0 >= sim->n) break;
      CUDA_OK(cudaStreamWaitEvent(sim->gstream[g], sim->ev_begin, 0));
      int nl = 0;
      int* mg = sim->mig ? sim->mig + (size_t)g * (nsub + 1) * (1 + LCR_MIGCAP) : nullptr;
      cudaStream_t* ms = sim->mig ? &sim->mstream[(size_t)g * (nsub + 1)] : nullptr;
      cudaEvent_t *mf = sim->mig ? &sim->mev_fork[(size_t)g * (nsub + 1)] : nullptr, *mj = sim->mig ? &sim->mev_join[(size_t)g * (nsub + 1)] : nullptr;
      int* er = sim->mig ? sim->early + (size_t)g * sim->early_stride : nullptr;
      int* jq = sim->jobq ? sim->jobq + (size_t)g * sim->jobq_stride : nullptr;
      if (jq) CUDA_OK(cudaMemsetAsync(jq, 0, sizeof(int) * LCR_JOBQ_HEADER, sim->gstream[g]));  // (k_ph_col clears the counters after every substep; this covers a step that was cut short)
      if (f32) LCR_DISPATCH_RET(float, sim->ncube, nl, step_phased(sim->cfg.n_substeps, sim->f.dm, sim->f.verts, sim->f.s, sim->f.gws, io, redo, env0, cnt, sim->perm, sim->gstream[g], mg, ms, mf, mj, jq, er));
      else LCR_DISPATCH_RET(double, sim->ncube, nl, step_phased(sim->cfg.n_substeps, sim->d.dm, sim->d.verts, sim->d.s, sim->d.gws, io, redo, env0, cnt, sim->perm, sim->gstream[g], mg, ms, mf, mj, jq, er));
      sim->launches += nl;
      CUDA_OK(cudaEventRecord(sim->ev_done[g], sim->gstream[g]));
      CUDA_OK(cudaStreamWaitEvent(st, sim->ev_done[g], 0));
      if (sim->mig)  // the BIG passes over the envs that migrated out of this chain
        for (int k = 0; k <= sim->cfg.n_substeps; k++) CUDA_OK(cudaStreamWaitEvent(st, mj[k], 0));  // (k == n_substeps: the begin stage)
    }
  } else if (sim->cfg.exec_mode == 2) {
    int W = 0;
    if (f32) LCR_DISPATCH_RET(float, sim->ncube, W, lockstep_warps(sim->ls_warps));
    else LCR_DISPATCH_RET(double, sim->ncube, W, lockstep_warps(sim->ls_warps));
    const int grid = (sim->n + W - 1) / W;
    if (sim->perm) {
      if (f32) lcr::Launch<float>::sched(sim->f.s, sim->perm, W, sim->ls_striped, sim->big, sim->tbig, st);
      else lcr::Launch<double>::sched(sim->d.s, sim->perm, W, sim->ls_striped, sim->big, sim->tbig, st);
      sim->launches++;
      predicted = true;
      start_big();
    }
    LCR_RUN(sim, step_lockstep(sim->f.dm, sim->f.verts, sim->f.s, io, redo, grid, W, W, sim->ls_flags, sim->perm, sim->prof, st),
            step_lockstep(sim->d.dm, sim->d.verts, sim->d.s, io, redo, grid, W, W, sim->ls_flags, sim->perm, sim->prof, st));
    sim->launches++;
  } else {
    LCR_RUN(sim, step(sim->f.dm, sim->f.verts, sim->f.s, io, redo, st), step(sim->d.dm, sim->d.verts, sim->d.s, io, redo, st));
    sim->launches++;
  }
  if (big_started) CUDA_OK(cudaStreamWaitEvent(st, sim->ev_big, 0));
  (void)predicted;
  // the envs that outgrew the fast workspace unexpectedly: the same step from the same start state over the big workspace
  if (env_int("LCR_NO_REDO", 0) == 0)  // (debug knob: timing of the main kernels alone; results are then wrong for those envs)
  LCR_RUN(sim, step_big(sim->f.dm, sim->f.verts, sim->f.s, io, redo, st), step_big(sim->d.dm, sim->d.verts, sim->d.s, io, redo, st));
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

int lcr_obs_dim(int task) { return (task == LCR_TASK_REACH || task == LCR_TASK_LIFT || task == LCR_TASK_PUSH_LOOP) ? 15 : 18; }
int lcr_action_dim(const LcrEnvCfg* cfg) { return (cfg->action_mode ? 3 : 5) + (cfg->block_gripper ? 0 : 1); }

int lcr_create(const LcrModel* model, const double* hull_verts, const LcrEnvCfg* cfg, int n_envs, int device, int precision, LcrSim** out) {
  if (!model || !cfg || !out || (!hull_verts && model->nvert > 0)) return fail("lcr_create: null argument");
  if (n_envs <= 0) return fail("lcr_create: n_envs must be positive");
  if (model->ncube < 1 || model->ncube > LCR_MAXCUBE || model->nmesh > LCR_MAXMESH || model->npair > LCR_MAXPAIR ||
      model->nwall < 0 || model->nwall > LCR_MAXWALL || model->nmesh + 1 + model->ncube + model->nwall > LCR_MAXGEOM)
    return fail("lcr_create: model exceeds compiled caps");
  if (model->task < LCR_TASK_REACH || model->task > LCR_TASK_PUSH_LOOP) return fail("lcr_create: unknown task");
  if ((model->task == LCR_TASK_PUSH_LOOP) != (model->nwall > 0) || (model->task == LCR_TASK_PUSH_LOOP && (model->ncube != 1 || model->nwall != LCR_MAXWALL)))
    return fail("lcr_create: static wall boxes are the four rails of the PushCubeLoop scene (one cube)");
  if (precision != LCR_F32 && precision != LCR_F64) return fail("lcr_create: precision must be LCR_F32 or LCR_F64");
  if (cfg->exec_mode < 0 || cfg->exec_mode > 3) return fail("lcr_create: exec_mode must be 0 (fused), 1 (phased), 2 (lockstep) or 3 (flow)");
  if (cfg->exec_mode == 3 && cfg->n_substeps < 1) return fail("lcr_create: the flow mode needs n_substeps >= 1 (its phase queues always run the first substep)");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("lcr_create: no such CUDA device");
  CUDA_OK(cudaSetDevice(device));
  LcrSim* s = new LcrSim();
  s->precision = precision; s->device = device; s->n = n_envs; s->ncube = scene_class(model->task, model->ncube); s->task = model->task;
  s->model = *model; s->cfg = *cfg;
  // (every failure below goes through lcr_destroy: nothing allocated so far is leaked)
#define LCR_CREATE_OK(call) do { if ((call) != cudaSuccess) { fail(std::string("lcr_create: ") + #call + " failed: " + cudaGetErrorString(cudaGetLastError())); lcr_destroy(s); return 1; } } while (0)
  if (precision == LCR_F32 ? s->f.create(*model, hull_verts, *cfg, n_envs) : s->d.create(*model, hull_verts, *cfg, n_envs)) { lcr_destroy(s); return 1; }
  s->launches = 1;
  LCR_CREATE_OK(cudaMalloc(&s->redo, sizeof(int) * ((size_t)n_envs + 1)));
  LCR_CREATE_OK(cudaMemset(s->redo, 0, sizeof(int)));
  LCR_CREATE_OK(cudaMalloc(&s->seed_buf, 32 * (size_t)n_envs));
  LCR_CREATE_OK(cudaMalloc(&s->big, sizeof(int) * ((size_t)n_envs + 1)));
  LCR_CREATE_OK(cudaMemset(s->big, 0, sizeof(int)));
  LCR_CREATE_OK(cudaStreamCreateWithFlags(&s->bstream, cudaStreamNonBlocking));
  LCR_CREATE_OK(cudaEventCreateWithFlags(&s->ev_sched, cudaEventDisableTiming));
  LCR_CREATE_OK(cudaEventCreateWithFlags(&s->ev_big, cudaEventDisableTiming));
  // rows of the previous step from which an env starts on the big workspace, beside the main kernels, instead of being found out and
  // redone after them.  Phased chain: from the fast cap on (small CTAs, the BIG CTAs run among them: PushCube 16 384, 21.5 -> 20.9 ms per
  // step); lockstep: never (its CTAs take a whole SM each, the BIG CTAs would only wait for one)
  s->tbig = env_int("LCR_TBIG", cfg->exec_mode == 1 ? 97 : 100000);
  if (cfg->exec_mode == 1) {
    // measured on B200 in the stationary window (profiles/r02b_sweep_groups.txt): the chains of the groups overlap each other's launch
    // tails; ms per step with 1 / 2 / 4 groups: PushCube 16 384 15.9 / 15.2 / 14.9, StackTwoCubes 8 192 - / 13.5 / 12.6, PickPlace-ee
    // 8 192 - / 16.4 / 15.9, ReachCube 4 096 - / 7.6 / 7.4; more groups only shrink the launches (8: +4 %, 16: +10 % at 16 384 envs)
    int g = env_int("LCR_GROUPS", n_envs <= 16384 ? 4 : 2);
    g = std::max(1, std::min(16, std::min(g, n_envs)));
    for (int k = 0; k < g; k++) {
      LCR_CREATE_OK(cudaStreamCreateWithFlags(&s->gstream[k], cudaStreamNonBlocking));
      LCR_CREATE_OK(cudaEventCreateWithFlags(&s->ev_done[k], cudaEventDisableTiming));
      s->ngroups = k + 1;
    }
    LCR_CREATE_OK(cudaEventCreateWithFlags(&s->ev_begin, cudaEventDisableTiming));
    // work-aware seats (LCR_PH_SORT=0: identity): the envs with the most constraint rows in their previous step are launched
    // first in every phase kernel, so that the long Newton solves / MPR jobs do not end up in the tail of the launch
    if (env_int("LCR_PH_SORT", 1) != 0) LCR_CREATE_OK(cudaMalloc(&s->perm, sizeof(int) * (2 * (size_t)n_envs + 16)));
    if (env_int("LCR_PH_MIGRATE", 1) != 0) {
      const int nsub1 = std::max(1, cfg->n_substeps) + 1;  // one side stream per substep, and one for the begin stage
      const int nb = s->ngroups * nsub1;
      LCR_CREATE_OK(cudaMalloc(&s->mig, sizeof(int) * (size_t)nb * (1 + LCR_MIGCAP)));
      s->early_stride = (((n_envs + s->ngroups - 1) / s->ngroups) + 1 + 31) & ~31;
      LCR_CREATE_OK(cudaMalloc(&s->early, sizeof(int) * (size_t)s->early_stride * s->ngroups));
      s->mstream.assign(nb, nullptr); s->mev_fork.assign(nb, nullptr); s->mev_join.assign(nb, nullptr);
      for (int k = 0; k < nb; k++) {
        LCR_CREATE_OK(cudaStreamCreateWithFlags(&s->mstream[k], cudaStreamNonBlocking));
        LCR_CREATE_OK(cudaEventCreateWithFlags(&s->mev_fork[k], cudaEventDisableTiming));
        LCR_CREATE_OK(cudaEventCreateWithFlags(&s->mev_join[k], cudaEventDisableTiming));
      }
    }
    // (experiment knob, off: one narrowphase job queue per chain, drained by a fixed grid of warps through an atomic ticket, instead of 8
    // warps per env.  Measured on B200: PushCube 16 384 15.04 -> 14.93 ms per step, ReachCube 4 096 7.34 -> 7.85, StackTwoCubes 8 192
    // 12.6 -> 13.0 -- the jobs of one env scatter over the SMs and lose the locality of its workspace and hull vertices)
    if (env_int("LCR_PH_JOBQ", 0) != 0) {
      const int per = (n_envs + s->ngroups - 1) / s->ngroups;
      int words = 0;
      if (precision == LCR_F32) LCR_DISPATCH_RET(float, s->ncube, words, jobq_words(per));
      else LCR_DISPATCH_RET(double, s->ncube, words, jobq_words(per));
      s->jobq_stride = (words + 31) & ~31;
      LCR_CREATE_OK(cudaMalloc(&s->jobq, sizeof(int) * (size_t)s->jobq_stride * s->ngroups));
      LCR_CREATE_OK(cudaMemset(s->jobq, 0, sizeof(int) * (size_t)s->jobq_stride * s->ngroups));
    }
    s->ph_striped = env_int("LCR_PH_STRIPED", 1);  // 1: ranks dealt out to the groups like cards; 0: group 0 holds the heaviest envs
    s->use_graph = env_int("LCR_GRAPH", 1);
    if (s->use_graph) {
      LCR_CREATE_OK(cudaStreamCreateWithFlags(&s->cstream, cudaStreamNonBlocking));
      LCR_CREATE_OK(cudaMalloc(&s->act_buf, sizeof(float) * (size_t)n_envs * lcr_action_dim(cfg)));
    }
  }
  if (cfg->exec_mode == 2) {  // tuning overrides for experiments; the defaults are the measured best
    s->ls_warps = env_int("LCR_LS_WARPS", 0);
    s->ls_flags = env_int("LCR_LS_FLAGS", 23);
    const int srt = env_int("LCR_LS_SORT", -1);
    if (srt != 0) {
      LCR_CREATE_OK(cudaMalloc(&s->perm, sizeof(int) * ((size_t)n_envs + 16)));
      // striped seats while the step time is set by the most expensive env, sorted seats once there are many waves
      s->ls_striped = srt == 2 ? 0 : (srt == 1 ? 1 : (n_envs < 12288 ? 1 : 0));
    }
  }
  if (cfg->exec_mode == 3 && create_flow(s)) { lcr_destroy(s); return 1; }
#undef LCR_CREATE_OK
  *out = s;
  return 0;
}

int lcr_destroy(LcrSim* sim) {
  if (!sim) return 0;
  cudaSetDevice(sim->device);
  if (sim->precision == LCR_F32) sim->f.destroy(); else sim->d.destroy();
  for (int k = 0; k < 16; k++) {
    if (sim->gstream[k]) cudaStreamDestroy(sim->gstream[k]);
    if (sim->ev_done[k]) cudaEventDestroy(sim->ev_done[k]);
  }
  if (sim->ev_begin) cudaEventDestroy(sim->ev_begin);
  for (cudaStream_t x : sim->mstream) if (x) cudaStreamDestroy(x);
  for (cudaEvent_t x : sim->mev_fork) if (x) cudaEventDestroy(x);
  for (cudaEvent_t x : sim->mev_join) if (x) cudaEventDestroy(x);
  cudaFree(sim->mig);
  cudaFree(sim->early);
  cudaFree(sim->jobq);
  if (sim->gexec) cudaGraphExecDestroy(sim->gexec);
  if (sim->cstream) cudaStreamDestroy(sim->cstream);
  cudaFree(sim->act_buf);
  if (sim->bstream) cudaStreamDestroy(sim->bstream);
  if (sim->ev_sched) cudaEventDestroy(sim->ev_sched);
  if (sim->ev_big) cudaEventDestroy(sim->ev_big);
  cudaFree(sim->big);
  cudaFree(sim->perm); cudaFree(sim->redo); cudaFree(sim->fq_mem); cudaFree(sim->fq_rings); cudaFree(sim->seed_buf);
  delete sim;
  return 0;
}

int lcr_seed(LcrSim* sim, const uint64_t* h_state, const uint8_t* d_mask, void* stream) {
  WITH_DEVICE(sim);
  if (!h_state) return fail("lcr_seed: null state");
  // (pageable host memory: the copy is staged by the runtime before the call returns, the caller's buffer is free afterwards)
  CUDA_OK(cudaMemcpyAsync(sim->seed_buf, h_state, 32 * (size_t)sim->n, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  if (sim->precision == LCR_F32) lcr::Launch<float>::seed(sim->f.s, sim->seed_buf, d_mask, (cudaStream_t)stream);
  else lcr::Launch<double>::seed(sim->d.s, sim->seed_buf, d_mask, (cudaStream_t)stream);
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_reset(LcrSim* sim, const uint8_t* d_mask, float* d_obs, void* stream) {
  WITH_DEVICE(sim);
  LCR_RUN(sim, reset(sim->f.dm, sim->f.verts, sim->f.s, d_mask, d_obs, (cudaStream_t)stream), reset(sim->d.dm, sim->d.verts, sim->d.s, d_mask, d_obs, (cudaStream_t)stream));
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_step_rec(LcrSim* sim, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated,
                 uint8_t* d_success, float* d_record, void* stream) {
  WITH_DEVICE(sim);
  if (!d_actions || !d_obs || !d_reward || !d_terminated || !d_truncated || !d_success) return fail("lcr_step: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const StepIO io{d_actions, d_obs, d_reward, d_terminated, d_truncated, d_success, d_record};
  if (sim->cfg.exec_mode == 3) {
    // one scheduler launch + one persistent kernel; the envs that need the big workspace are handled inside
    LCR_RUN(sim, step_flow(sim->f.dm, sim->f.verts, sim->f.s, sim->f.gws, io, &sim->fq, sim->flow_grid, sim->flow_bigcta, sim->flow_flags, sim->flow_thi,
                           sim->flow_tbig, sim->flow_stats, st),
            step_flow(sim->d.dm, sim->d.verts, sim->d.s, sim->d.gws, io, &sim->fq, sim->flow_grid, sim->flow_bigcta, sim->flow_flags, sim->flow_thi,
                      sim->flow_tbig, sim->flow_stats, st));
    sim->launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (sim->cfg.exec_mode == 1 && sim->use_graph) {
    // phased chain as one CUDA graph launch: captured on an internal stream (the caller's may be the legacy default stream,
    // which cannot be captured), replayed on the caller's stream
    CUDA_OK(cudaMemcpyAsync(sim->act_buf, d_actions, sizeof(float) * (size_t)sim->n * lcr_action_dim(&sim->cfg), cudaMemcpyDeviceToDevice, st));
    StepIO gio = io;
    gio.actions = sim->act_buf;
    const void* key[6] = {d_obs, d_reward, d_terminated, d_truncated, d_success, d_record};
    if (!sim->gexec || std::memcmp(key, sim->gkey, sizeof key) != 0) {
      if (sim->gexec) { cudaGraphExecDestroy(sim->gexec); sim->gexec = nullptr; }
      const int l0 = sim->launches;
      CUDA_OK(cudaStreamBeginCapture(sim->cstream, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_step(sim, gio, sim->cstream);
      cudaGraph_t graph = nullptr;
      const cudaError_t ec = cudaStreamEndCapture(sim->cstream, &graph);
      sim->graph_nodes = sim->launches - l0;
      sim->launches = l0;
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ec != cudaSuccess) return fail(std::string("lcr_step: graph capture failed: ") + cudaGetErrorString(ec));
      const cudaError_t ei = cudaGraphInstantiate(&sim->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ei != cudaSuccess) { sim->gexec = nullptr; return fail(std::string("lcr_step: cudaGraphInstantiate: ") + cudaGetErrorString(ei)); }
      std::memcpy(sim->gkey, key, sizeof key);
    }
    CUDA_OK(cudaGraphLaunch(sim->gexec, st));
    sim->launches += sim->graph_nodes;  // kernels launched by the replay
    return 0;
  }
  return enqueue_step(sim, io, st);
}

int lcr_step(LcrSim* sim, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated,
             uint8_t* d_success, void* stream) {
  return lcr_step_rec(sim, d_actions, d_obs, d_reward, d_terminated, d_truncated, d_success, nullptr, stream);
}

int lcr_get_state(LcrSim* sim, double* q, double* v, double* c, double* w, double* a, int32_t* i, uint64_t* rng, void* stream) {
  WITH_DEVICE(sim);
  const int ncu = sim->model.ncube;
  if (sim->precision == LCR_F32) lcr::Launch<float>::get_state(ncu, sim->f.s, q, v, c, w, a, i, (unsigned long long*)rng, (cudaStream_t)stream);
  else lcr::Launch<double>::get_state(ncu, sim->d.s, q, v, c, w, a, i, (unsigned long long*)rng, (cudaStream_t)stream);
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_set_state(LcrSim* sim, const double* q, const double* v, const double* c, const double* w, const double* a, const int32_t* i,
                  const uint64_t* rng, void* stream) {
  WITH_DEVICE(sim);
  const int ncu = sim->model.ncube;
  if (sim->precision == LCR_F32) lcr::Launch<float>::set_state(ncu, sim->f.s, q, v, c, w, a, i, (const unsigned long long*)rng, (cudaStream_t)stream);
  else lcr::Launch<double>::set_state(ncu, sim->d.s, q, v, c, w, a, i, (const unsigned long long*)rng, (cudaStream_t)stream);
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_substeps(LcrSim* sim, int n, void* stream) {
  WITH_DEVICE(sim);
  if (n < 0) return fail("lcr_substeps: n < 0");
  cudaStream_t st = (cudaStream_t)stream;
  const Redo redo = redo_of(sim);
  CUDA_OK(cudaMemsetAsync(sim->redo, 0, sizeof(int), st));
  LCR_RUN(sim, substeps(sim->f.dm, sim->f.verts, sim->f.s, n, redo, st), substeps(sim->d.dm, sim->d.verts, sim->d.s, n, redo, st));
  LCR_RUN(sim, substeps_big(sim->f.dm, sim->f.verts, sim->f.s, n, redo, st), substeps_big(sim->d.dm, sim->d.verts, sim->d.s, n, redo, st));
  sim->launches += 2;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_ik(LcrSim* sim, const float* d_ee_target, float* d_q_out, void* stream) {
  WITH_DEVICE(sim);
  if (!d_ee_target || !d_q_out) return fail("lcr_ik: null buffer");
  LCR_RUN(sim, ik(sim->f.dm, sim->f.verts, sim->f.s, d_ee_target, d_q_out, (cudaStream_t)stream), ik(sim->d.dm, sim->d.verts, sim->d.s, d_ee_target, d_q_out, (cudaStream_t)stream));
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_get_diag(LcrSim* sim, int32_t* d_diag, void* stream) {
  WITH_DEVICE(sim);
  if (!d_diag) return fail("lcr_get_diag: null buffer");
  if (sim->precision == LCR_F32) lcr::Launch<float>::get_diag(sim->f.s, d_diag, (cudaStream_t)stream);
  else lcr::Launch<double>::get_diag(sim->d.s, d_diag, (cudaStream_t)stream);
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_debug_contacts(LcrSim* sim, double* d_contacts, int32_t* d_ncon, void* stream) {
  WITH_DEVICE(sim);
  if (!d_contacts || !d_ncon) return fail("lcr_debug_contacts: null buffer");
  LCR_RUN(sim, debug_contacts(sim->f.dm, sim->f.verts, sim->f.s, d_contacts, d_ncon, (cudaStream_t)stream),
          debug_contacts(sim->d.dm, sim->d.verts, sim->d.s, d_contacts, d_ncon, (cudaStream_t)stream));
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_pack_outputs(LcrSim* sim, const float* d_obs, const float* d_reward, const uint8_t* d_terminated, const uint8_t* d_truncated,
                     const uint8_t* d_success, float* d_record, void* stream) {
  WITH_DEVICE(sim);
  if (!d_obs || !d_reward || !d_terminated || !d_truncated || !d_success || !d_record) return fail("lcr_pack_outputs: null buffer");
  lcr::Launch<float>::pack(d_obs, d_reward, d_terminated, d_truncated, d_success, d_record, sim->n, lcr_obs_dim(sim->task), (cudaStream_t)stream);
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_pose_slots(const LcrSim* sim) {
  if (!sim) return 0;
  int n = 0;
  LCR_DISPATCH_RET(float, sim->ncube, n, pose_slots());
  return n;
}

int lcr_body_poses(LcrSim* sim, float* d_poses, void* stream) {
  WITH_DEVICE(sim);
  if (!d_poses) return fail("lcr_body_poses: null buffer");
  LCR_RUN(sim, poses(sim->f.dm, sim->f.s, d_poses, (cudaStream_t)stream), poses(sim->d.dm, sim->d.s, d_poses, (cudaStream_t)stream));
  sim->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_record_append(const float* d_obs, int obs_dim, const float* d_actions, int action_dim, const uint8_t* d_terminated, const uint8_t* d_truncated,
                      int n_envs, int horizon, float* d_traj, int32_t* d_len, float* d_pool, int32_t* d_pool_meta, int32_t* d_count, int pool_cap,
                      void* stream) {
  if (!d_obs || !d_actions || !d_terminated || !d_truncated || !d_traj || !d_len || !d_pool || !d_pool_meta || !d_count)
    return fail("lcr_record_append: null buffer");
  if (obs_dim < 12 || action_dim < 1 || 12 + action_dim > 32 || n_envs <= 0 || horizon <= 0 || pool_cap <= 0)
    return fail("lcr_record_append: obs_dim >= 12, 1 <= action_dim <= 20, positive n_envs / horizon / pool_cap required");
  lcr::Launch<float>::rec_append(d_obs, obs_dim, d_actions, action_dim, d_terminated, d_truncated, n_envs, horizon, d_traj, d_len, d_pool, d_pool_meta,
                                 d_count, pool_cap, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int lcr_debug_phase_clocks(LcrSim* sim, long long* d_clocks) {
  if (!sim) return fail("null handle");
  if (sim->cfg.exec_mode != 2) return fail("lcr_debug_phase_clocks: lockstep mode only");
  sim->prof = d_clocks;
  return 0;
}

int lcr_debug_flow_stats(LcrSim* sim, unsigned long long* d_stats) {
  if (!sim) return fail("null handle");
  if (sim->cfg.exec_mode != 3) return fail("lcr_debug_flow_stats: flow mode only");
  sim->flow_stats = d_stats;
  return 0;
}

int lcr_flow_status(LcrSim* sim, int32_t* h_status) {
  WITH_DEVICE(sim);
  if (!h_status) return fail("lcr_flow_status: null buffer");
  for (int k = 0; k < 8; k++) h_status[k] = 0;
  if (sim->cfg.exec_mode != 3) return 0;
  unsigned v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CUDA_OK(cudaMemcpy(&v[0], sim->fq.ctl + 64 * LCR_FQ_NQ, sizeof(unsigned), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(&v[1], sim->fq.ctl + 64 * LCR_FQ_NQ + 32, 7 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 8; k++) h_status[k] = (int32_t)v[k];
  return 0;
}

int lcr_n_envs(const LcrSim* sim) { return sim ? sim->n : 0; }
int lcr_kernel_launches(const LcrSim* sim) { return sim ? sim->launches : 0; }
const char* lcr_last_error(void) { return g_err.c_str(); }
const char* lcr_version(void) { return "lcrsim 0.2 (sm_100a)"; }
int lcr_sizeof_model(void) { return (int)sizeof(LcrModel); }
int lcr_sizeof_cfg(void) { return (int)sizeof(LcrEnvCfg); }

}  // extern "C"
