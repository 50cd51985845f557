// lcr_device.cuh -- device-side data structures of liblcrsim (sm_100a).
//
// Execution model: ONE WARP PER ENVIRONMENT, one 32-thread CTA per warp (envs finish at different
// times -- solver iterations and contact counts are data dependent -- so warps must not wait for
// each other; the hardware CTA scheduler does the load balancing).  The persistent state of an env
// is one contiguous, 16-byte aligned record in HBM; the warp stages it into shared memory with
// coalesced 128-bit loads (lane l moves uint4 #l of the record), runs the whole control step
// (action map / IK -> n_substeps x mj_step -> observation / reward) out of its private
// shared-memory workspace, and writes the record back the same way.  Small dense algebra is done
// with lanes <-> rows/columns and warp shuffles; there is no tensor-core work on this path.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/lcr_model.h"
#include "../../include/lcrsim.h"

#define LCR_WPB 1          // warps (= envs) per CTA: every env is an independent 32-thread CTA
#define FULLMASK 0xffffffffu

// Scene class = the template parameter NC of the workspace and of every kernel: 1 / 2 = that many free cubes and no
// static boxes (Reach / Push / Lift / PickPlace, StackTwoCubes); LCR_NC_LOOP = one cube + the LCR_MAXWALL static wall
// boxes of PushCubeLoop.  Boxes are indexed cubes first, walls after them; box c has the pose slot LCR_NABODY + c of
// Ws::xpos / xquat / xmat (the wall slots are constants rewritten by every kinematics pass).
// Bit 4 of NC (LCR_NC_BIG) selects the BIG workspace caps (LCR_MAXCON_BIG / LCR_MAXEFC_BIG, include/lcr_model.h): the
// same device functions instantiated over a larger workspace, used by the slow path that takes over the envs whose
// contact list or constraint rows outgrow the fast caps.
// phased chain: envs whose constraint rows outgrow the fast workspace in substep k of chain g are appended to the migration list (g, k)
// (1 + LCR_MIGCAP ints: count, then env | k << 24) and resume over the big workspace at once; beyond the cap they take the redo pass
#define LCR_MIGCAP 64
// phased chain: narrowphase job queue of one chain (ints): entry count, next ticket (own cache lines), then the entries env << 6 | candidate
#define LCR_JOBQ_COUNT 0
#define LCR_JOBQ_NEXT 32
#define LCR_JOBQ_ITEMS 64
#define LCR_JOBQ_HEADER LCR_JOBQ_ITEMS
#define LCR_NC_LOOP 5
#define LCR_NC_BIG 16
template <int NC> struct Scene {
  static constexpr int S = NC & 15;
  static constexpr bool BIG = (NC & LCR_NC_BIG) != 0;
  static constexpr int NCUBE = S == LCR_NC_LOOP ? 1 : S, NWALL = S == LCR_NC_LOOP ? LCR_MAXWALL : 0, NBOX = NCUBE + NWALL;
  static constexpr int MAXCON = BIG ? LCR_MAXCON_BIG : LCR_MAXCON, MAXEFC = BIG ? LCR_MAXEFC_BIG : LCR_MAXEFC;
  static constexpr int MAXCAND = 3 * MAXEFC / 8;  // candidate results (8 words each) live in the e_w / e_g / e_p rows
};
// body id of box c in contact records: the cube's free body, or the world (-1) for a static wall
template <int NC> __host__ __device__ constexpr int box_body(int c) { return c < Scene<NC>::NCUBE ? LCR_NABODY + c : -1; }
inline int scene_class(int task, int ncube) { return task == LCR_TASK_PUSH_LOOP ? LCR_NC_LOOP : ncube; }

// contact parameter classes, precomputed on the host by the MuJoCo mixing rule
template <typename T>
struct CPar {
  T fr[5];      // friction: slide, slide, spin, roll, roll
  T K, B;       // reference stiffness / damping from solref (refsafe) and clamped dmax
  T si[5];      // clamped solimp
  int dim;
};

template <typename T>
struct DevModel {
  int task, ncube, nq, nv, nmesh, nvert, npair, site_body, iterations, ls_iterations;
  T timestep, impratio, tolerance, ls_tolerance, meaninertia, gravity[3];
  T body_pos[LCR_NABODY][3], body_quat[LCR_NABODY][4], body_ipos[LCR_NABODY][3], body_iquat[LCR_NABODY][4];
  T body_mass[LCR_NABODY], body_inertia[LCR_NABODY][3], body_invweight0[LCR_NABODY + LCR_MAXCUBE][2];
  T jnt_axis[LCR_NARM][3], jnt_range[LCR_NARM][2], jnt_armature[LCR_NARM], jnt_damping[LCR_NARM];
  T jnt_frcrange[LCR_NARM][2], dof_invweight0[LCR_NARM], act_kp[LCR_NARM], act_kv[LCR_NARM], act_ctrlrange[LCR_NARM][2];
  T site_pos[3];
  T cube_mass[LCR_MAXCUBE], cube_inertia[LCR_MAXCUBE], cube_qpos0[LCR_MAXCUBE][3];
  T cube_size[LCR_MAXBOX][3];  // half sizes of the boxes: cubes, then the static walls
  T wall_pos[LCR_MAXWALL][3];
  int mesh_body[LCR_MAXMESH], mesh_vertadr[LCR_MAXMESH], mesh_vertnum[LCR_MAXMESH];
  T mesh_center[LCR_MAXMESH][3], mesh_half[LCR_MAXMESH][3], mesh_rbound[LCR_MAXMESH], mesh_com[LCR_MAXMESH][3];
  int pair_g1[LCR_MAXPAIR], pair_g2[LCR_MAXPAIR];
  // contact parameter classes (one contiguous table: Ws::c_par stores indices relative to par_limit)
  CPar<T> par_limit[LCR_NARM];
  CPar<T> par_floor_cube[LCR_MAXCUBE];
  CPar<T> par_cube_cube;
  CPar<T> par_floor_mesh[LCR_MAXMESH];
  CPar<T> par_cube_mesh[LCR_MAXBOX][LCR_MAXMESH];  // box (cube or wall) vs mesh
  CPar<T> par_wall_cube[LCR_MAXWALL][LCR_MAXCUBE];
  CPar<T> par_mesh_mesh[LCR_MAXPAIR];
  // env config
  int action_mode, block_gripper, reward_type, n_substeps, max_episode_steps, autoreset, collision_mask;
  T distance_threshold, height_threshold;
  double cube_low[3], cube_high[3], target_low[3], target_high[3];  // reset draws are float64 like numpy
  // PushCubeLoop: goal region centres and the sampling / overlap half box (push_cube_loop_env.py:127-135), float64 like numpy
  double goal_center[2][3], goal_high[3];
  __host__ __device__ const CPar<T>* par(int idx) const { return par_limit + idx; }
  __host__ __device__ int par_index(const CPar<T>* p) const { return (int)(p - par_limit); }
};

// Global (HBM) state: one record per env.  `st` [n][NFP] of T, fields qpos[nq] | qvel[nv] | ctrl[6] |
// warm[nv] | aux[LCR_NAUX] (time, target[3], site_xpos[3], cube_xpos[6]) | pad to a multiple of 16 B.
// `ib` [n][16] int32: elapsed, needs_reset | diag[6] | PCG64 state_hi, state_lo, inc_hi, inc_lo (4 x u64).
// `sa` [n][SA_BYTES]: separating-axis cache dir[16][3], S[16][2], u[16][2][3] of T | key[16] int16 | next | pad.
#define LCR_IB_WORDS 16
template <typename T>
struct DevState {
  T* st;
  int32_t* ib;
  unsigned char* sa;  // [n][Ws::SA_BYTES] separating-axis cache (performance only; emptied by init / set_state)
  int n, nfp;
};

template <typename T, int NC>
struct Ws {  // per-warp shared-memory workspace
  static constexpr int NCU = Scene<NC>::NCUBE;
  static constexpr int MAXCON = Scene<NC>::MAXCON, MAXEFC = Scene<NC>::MAXEFC, MAXCAND = Scene<NC>::MAXCAND;
  static constexpr int NQ = LCR_NARM + 7 * NCU, NVV = LCR_NARM + 6 * NCU, NB = LCR_NABODY + Scene<NC>::NBOX;
  static constexpr int NF = NQ + 2 * NVV + LCR_NARM + LCR_NAUX;
  static constexpr int JS = NVV + 1;  // padded row stride of J (odd -> conflict-free row-parallel access)
  static constexpr int NFP = (NF + 3) & ~3;  // record length in T, multiple of 16 bytes
  alignas(16) T st[NFP];
  alignas(16) int ints[LCR_NINT];  // ints | diag | rng are contiguous = the 64-byte `ib` record
  int diag[LCR_NDIAG];
  unsigned long long rng[4];
  // kinematics
  alignas(16) T xpos[NB][3];  // (16-byte aligned starts: the phased kernels stage the workspace by region, see load_ws_range)
  T xquat[NB][4], xmat[NB][9], xipos[LCR_NABODY][3], axis[LCR_NARM][3];
  T Iw[LCR_NABODY][6];
  T gc[LCR_MAXMESH][3];  // world centres of the mesh bounding spheres / boxes
  alignas(16) T M[LCR_NARM][LCR_NARM];
  T bias[NVV], smooth[NVV], qacc_smooth[NVV], qacc[NVV], Ma[NVV], grad[NVV], search[NVV], Mv[NVV];
  alignas(16) T H[NVV][NVV + 1];
  // contacts
  alignas(16) T c_pos[MAXCON][3];
  T c_frame[MAXCON][9], c_dist[MAXCON], c_mu[MAXCON], c_c1[MAXCON], c_c2[MAXCON];
  short c_par[MAXCON];  // index into the contiguous CPar tables of DevModel, see DevModel::par()
  short c_efc[MAXCON];
  signed char c_b1[MAXCON], c_b2[MAXCON];
  // constraint rows
  alignas(16) T e_pos[MAXEFC];
  T e_D[MAXEFC], e_aref[MAXEFC], e_jar[MAXEFC];
  union {  // solver rows / RNE temporaries of inertia_and_bias (dead before the constraint rows are built)
    struct { T e_jv[MAXEFC], e_force[MAXEFC]; };
    struct { T rw[LCR_NABODY][3], ral[LCR_NABODY][3], ra[LCR_NABODY][3], F[LCR_NABODY][3], Nn[LCR_NABODY][3]; };
  };
  T e_w[MAXEFC], e_g[MAXEFC], e_p[MAXEFC];  // Hessian pieces, see contact_eval
  alignas(16) short e_unit[MAXEFC];  // >= 0: contact index; < 0: limit row of joint -1-e_unit
  signed char e_r[MAXEFC];  // row index within its contact
  alignas(16) int ncon;
  int nefc, nlim;
  // separating-axis cache of the convex narrowphase (see lcr_convex.cuh): one contiguous 16-byte aligned block
  // that travels with the env record (DevState::sa)
  alignas(16) T sa_dir[LCR_NSA][3];
  T sa_S[LCR_NSA][2];     // per hull of the pair: support value about the hull's bounding-sphere centre, along +-axis ...
  T sa_u[LCR_NSA][2][3];  // ... and that direction in the hull's body frame, both at the last exact evaluation
  short sa_key[LCR_NSA];
  int sa_next;
  int sa_pad[3];
  static constexpr int SA_WORDS = LCR_NSA * 11;  // dir, S, u
  static constexpr int SA_BYTES = SA_WORDS * (int)sizeof(T) + LCR_NSA * 2 + 16;
  alignas(16) short cand_key[MAXCAND];  // convex-pair candidates of this substep (results alias e_w / e_g / e_p)
  int ncand;
  int skip;          // phased execution: this env was auto-reset by the current step, substep kernels pass
  int redo_forward;  // phased execution: state was reset after a bad qacc, re-run mj_forward before integrating
  int ovf;           // a cap of THIS workspace was hit since the record was loaded (fast path: the env moves to the BIG path)
  int substep;       // flow execution: mj_step calls of the current env.step completed so far
  int jobs_left;     // flow execution: narrowphase jobs of the current substep still running (global atomic counter)
  alignas(16) T J[MAXEFC][JS];  // last member: only the first nefc rows are live (and staged)

  __device__ T* qpos() { return st; }
  __device__ T* qvel() { return st + NQ; }
  __device__ T* ctrl() { return st + NQ + NVV; }
  __device__ T* warm() { return st + NQ + NVV + LCR_NARM; }
  __device__ T* aux() { return st + NQ + 2 * NVV + LCR_NARM; }
  __device__ T* target() { return aux() + 1; }
  __device__ T* site_xpos() { return aux() + 4; }
  __device__ T* cube_xpos(int c) { return aux() + 7 + 3 * c; }
};

// Device buffers of one lcr_step call (borrowed from the caller, row-major over the envs).  `rec` (optional) is the packed
// float32 record [n][obs_dim + 4] = obs | reward | terminated | truncated | success: the unit of the one all-gather per step
// and of the device -> host copy, written by the step kernels themselves (no separate pack launch).
struct StepIO {
  const float* actions;
  float* obs;
  float* reward;
  uint8_t* term;
  uint8_t* trunc;
  uint8_t* succ;
  float* rec;
};
// Envs that hit a cap of the FAST workspace during a step are not written back; they are appended here and the same
// step is redone from the same start state over the BIG workspace (k_step / k_substeps with NC | LCR_NC_BIG).
struct Redo {
  int* count;  // null: no redirection (BIG kernels: whatever exceeds the BIG caps is dropped and counted in diag)
  int* list;
};
// Work queues of the flow kernel (lcr_flow.cuh): one ring per phase and priority.
enum { FQ_BEGIN = 0, FQ_DYN = 1, FQ_JOB = 2, FQ_COL = 3, FQ_SOL = 4, FQ_END = 5, FQ_BIG = 6, FQ_NPH = 7 };
#define LCR_FQ_NQ (2 * FQ_NPH)  // two priorities per phase: queue 2 p = express (envs that were expensive in their previous step)
#define LCR_FQ_EMPTY 0xffffffffu
#define LCR_FQ_CTL_WORDS (64 * LCR_FQ_NQ + 64 + 64)
#define LCR_FQ_MAXENV (1 << 20)  // item = env (20 bits) | aux (10 bits) << 20 | express << 30
struct FlowQ {
  unsigned* ctl;  // [LCR_FQ_NQ][64]: head at +0, tail at +32 (own 128-byte lines) | remaining | error -- zeroed by k_sched_flow
  unsigned* ring[LCR_FQ_NQ];  // entries are LCR_FQ_EMPTY except between a push and its pop
  unsigned mask[LCR_FQ_NQ];
};

// host-side launchers.  LaunchNC<T, S> (everything that depends on the scene class S = 1, 2, LCR_NC_LOOP) is instantiated
// in one translation unit per (T, S) (lcr_nc.cu); Launch<T> (scene-independent kernels) in lcr_common.cu.
namespace lcr {
template <typename T, int S>
struct LaunchNC {
  static void prepare();
  static size_t ws_bytes();
  static void reset(const DevModel<T>* dm, const T* verts, DevState<T> s, const uint8_t* mask, float* obs, cudaStream_t st);
  static void step(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo redo, cudaStream_t st);
  static void step_big(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo list, cudaStream_t st);
  static int lockstep_warps(int warps);
  static void step_lockstep(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo redo, int grid, int warps, int epc, int flags,
                            const int* perm, long long* prof, cudaStream_t st);
  static void step_big_resume(const DevModel<T>* dm, const T* verts, DevState<T> s, const void* gws, StepIO io, const int* mig, cudaStream_t st);
  static int step_phased(int n_substeps, const DevModel<T>* dm, const T* verts, DevState<T> s, void* gws, StepIO io, Redo redo, int env0, int cnt,
                         const int* perm, cudaStream_t st, int* mig, cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join, int* jobq, int* early);
  static int jobq_words(int cnt);
  static int flow_warps();
  static int flow_bigslots();
  static int flow_smem();
  static void step_flow(const DevModel<T>* dm, const T* verts, DevState<T> s, void* gws, StepIO io, const void* fq, int grid, int nbigcta, int flags,
                        int t_hi, int t_big, unsigned long long* stats, cudaStream_t st);
  static void substeps(const DevModel<T>* dm, const T* verts, DevState<T> s, int n, Redo redo, cudaStream_t st);
  static void substeps_big(const DevModel<T>* dm, const T* verts, DevState<T> s, int n, Redo list, cudaStream_t st);
  static void ik(const DevModel<T>* dm, const T* verts, DevState<T> s, const float* target, float* q_out, cudaStream_t st);
  static void debug_contacts(const DevModel<T>* dm, const T* verts, DevState<T> s, double* out, int32_t* ncon, cudaStream_t st);
  static int pose_slots();
  static void poses(const DevModel<T>* dm, DevState<T> s, float* out, cudaStream_t st);
};
template <typename T>
struct Launch {
  static void sched(DevState<T> s, int* perm, int W, int striped, int* big, int tbig, cudaStream_t st);
  static void pack(const float* obs, const float* reward, const uint8_t* term, const uint8_t* trunc, const uint8_t* succ, float* rec, int n, int od,
                   cudaStream_t st);
  static void rec_append(const float* obs, int od, const float* act, int A, const uint8_t* term, const uint8_t* trunc, int n, int h, float* traj,
                         int32_t* len, float* pool, int32_t* meta, int32_t* count, int cap, cudaStream_t st);
  static void get_state(int ncube, DevState<T> s, double* qpos, double* qvel, double* ctrl, double* warm, double* aux, int32_t* ints,
                        unsigned long long* rng, cudaStream_t st);
  static void set_state(int ncube, DevState<T> s, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                        const double* aux, const int32_t* ints, const unsigned long long* rng, cudaStream_t st);
  static void init_state(const DevModel<T>* dm, DevState<T> s, cudaStream_t st);
  static void get_diag(DevState<T> s, int32_t* out, cudaStream_t st);
  static void seed(DevState<T> s, const unsigned long long* d_state, const uint8_t* d_mask, cudaStream_t st);
};
}  // namespace lcr
