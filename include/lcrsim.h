/* lcrsim.h -- C-ABI of liblcrsim.so: the B200 batched replacement for the per-step hot path of
 * gym_lowcostrobot.envs.{ReachCube,PushCube,LiftCube,PickPlaceCube,StackTwoCubes,PushCubeLoop}Env.
 *
 * Every entry point below replaces a piece of the reference's Python->MuJoCo interface; the
 * reference file:line it stands in for is cited per function.  All `d_*` pointers are BORROWED
 * DEVICE pointers (owned by the caller, e.g. torch tensors); the library never frees them and
 * never synchronises the host -- each call only enqueues work on `stream` (a cudaStream_t passed
 * as void*; NULL = the legacy default stream).  Calls return 0 on success, non-zero on error with
 * a message available from lcr_last_error().  One handle per device; a handle is not thread-safe,
 * distinct handles are independent.  No torch types appear in any signature.
 */
#ifndef LCRSIM_H_
#define LCRSIM_H_

#include <stdint.h>
#include "lcr_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LcrSim LcrSim;

enum { LCR_F32 = 0, LCR_F64 = 1 };

/* Observation record written by lcr_step/lcr_reset, row-major [n_envs][obs_dim] float32
 * (reference get_observation, reach_cube_env.py:281-295, push_cube_env.py:291-306,
 * stack_two_cubes_env.py:290-305; all `.astype(np.float32)`):
 *   Reach/Lift      : arm_qpos[6] arm_qvel[6] cube_pos[3]                      (15)
 *   Push/PickPlace  : arm_qpos[6] arm_qvel[6] target_pos[3] cube_pos[3]        (18)
 *   Stack           : arm_qpos[6] arm_qvel[6] cube_red_pos[3] cube_blue_pos[3] (18)
 *   PushCubeLoop    : arm_qpos[6] arm_qvel[6] cube_pos[3]                      (15)   (push_cube_loop_env.py:286-300) */
int lcr_obs_dim(int task);
/* Action width: {joint:5, ee:3} + (block_gripper ? 0 : 1)   (reach_cube_env.py:95-97). */
int lcr_action_dim(const LcrEnvCfg* cfg);

/* Replaces MjModel.from_xml_path + MjData(model) (reach_cube_env.py:89-90 and the same two lines
 * in the other four envs): uploads the compiled model constants and allocates the SoA state of
 * n_envs instances on `device`.  `hull_verts` is a HOST array [model->nvert][3] (body frame).
 * `precision` LCR_F32 is the product path; LCR_F64 runs the same kernels in double for
 * verification against the CPU oracle.  State is initialised like mj_resetData (qpos0, zeros). */
int lcr_create(const LcrModel* model, const double* hull_verts, const LcrEnvCfg* cfg, int n_envs, int device,
               int precision, LcrSim** out);
int lcr_destroy(LcrSim* sim);

/* Replaces Env.reset(seed=...)'s seeding (gymnasium: np_random = Generator(PCG64(SeedSequence(seed)))).
 * `h_state` is a HOST array [n_envs][4] of uint64 = (state_hi, state_lo, inc_hi, inc_lo) of each
 * env's numpy PCG64 bit generator; the device continues that exact stream for every later draw.
 * `d_mask` (DEVICE, [n_envs] uint8, may be NULL = all): only the selected envs are reseeded -- a masked reset with a
 * seed must not rewind the streams of the envs it does not reset.  `h_state` may be reused as soon as the call returns. */
int lcr_seed(LcrSim* sim, const uint64_t* h_state, const uint8_t* d_mask, void* stream);

/* Replaces Env.reset (reach_cube_env.py:297-311, push_cube_env.py:308-328, lift_cube_env.py:306-320,
 * pick_place_cube_env.py:316-336, stack_two_cubes_env.py:307-324, push_cube_loop_env.py:302-320): for every env with
 * d_mask[i] != 0 (all envs if d_mask is NULL) draw cube (and target / second cube) positions
 * from the env's PCG64 stream, write qpos, run mj_forward, and write the observation row.
 * Like the reference it does NOT reset qvel / ctrl / warmstart / time (nor PushCubeLoop's current goal, in whose
 * region the cube is drawn).  Rows of unmasked envs in d_obs are left untouched. */
int lcr_reset(LcrSim* sim, const uint8_t* d_mask, float* d_obs, void* stream);

/* Replaces Env.step (reach_cube_env.py:313-333 = apply_action :223-279 incl. inverse_kinematics
 * :148-221 and the 20x mj_step loop :276-277, get_observation :281-295, is_success/compute_reward
 * :335-348; and the same methods of the other four envs) plus TimeLimit (gym_lowcostrobot/__init__.py).
 * d_actions: [n_envs][action_dim] float32.  Outputs: d_obs [n_envs][obs_dim] f32, d_reward [n_envs]
 * f32, d_terminated / d_truncated / d_success [n_envs] uint8.
 * PushCubeLoop (push_cube_loop_env.py:322-383): reward = get_reward's overlap reward (+5 on success, overlap - 1, or the
 * clipped y distance to the goal edge), d_success = info["success"], the env's current goal switches on success,
 * d_terminated is always 0. */
int lcr_step(LcrSim* sim, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated,
             uint8_t* d_truncated, uint8_t* d_success, void* stream);
/* lcr_step that also writes the packed record of lcr_pack_outputs (d_record [n_envs][obs_dim + 4] float32, may be NULL)
 * from inside the step kernels: the 5-tuple of step() (reach_cube_env.py:333) as ONE buffer = the send buffer of the
 * one all-gather per step of the sharded path and the single device->host copy of a host-facing caller. */
int lcr_step_rec(LcrSim* sim, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated,
                 uint8_t* d_truncated, uint8_t* d_success, float* d_record, void* stream);

/* Packs the outputs of lcr_step into one float32 record per env, d_record [n_envs][obs_dim + 4] =
 * obs | reward | terminated | truncated | success (what the reference's step() returns as a 5-tuple,
 * reach_cube_env.py:333).  It is the send buffer of the one all-gather per step of the sharded multi-GPU path
 * and the single device->host copy of a host-facing caller. */
int lcr_pack_outputs(LcrSim* sim, const float* d_obs, const float* d_reward, const uint8_t* d_terminated,
                     const uint8_t* d_truncated, const uint8_t* d_success, float* d_record, void* stream);

/* Replaces direct reads/writes of MjData the reference performs (data.qpos / data.qvel / data.ctrl,
 * reach_cube_env.py:182,185,252,273,285-294,305-306) and provides checkpoint/resume.  Row-major
 * DEVICE arrays of float64 regardless of precision: qpos [n][nq], qvel [n][nv], ctrl [n][6],
 * warm (qacc_warmstart) [n][nv], aux [n][LCR_NAUX] = time, target[3], site_xpos[3],
 * cube_xpos[3*2]; ints [n][LCR_NINT] = elapsed_steps, needs_reset; rng [n][4] uint64 = the PCG64 state of the env's reset
 * stream (np_random of the reference env, state_hi, state_lo, inc_hi, inc_lo) -- without it a restored rollout would draw
 * other cube / target positions at its next reset than the uninterrupted one.  Any pointer may be NULL.
 * PushCubeLoop keeps `current_goal` (0 / 1, push_cube_loop_env.py:136) in target[0]; aux[0] is info["timestamp"]. */
#define LCR_NAUX 13
#define LCR_NINT 2
int lcr_get_state(LcrSim* sim, double* d_qpos, double* d_qvel, double* d_ctrl, double* d_warm, double* d_aux,
                  int32_t* d_ints, uint64_t* d_rng, void* stream);
int lcr_set_state(LcrSim* sim, const double* d_qpos, const double* d_qvel, const double* d_ctrl,
                  const double* d_warm, const double* d_aux, const int32_t* d_ints, const uint64_t* d_rng, void* stream);

/* Advance `n` raw substeps (mujoco.mj_step, reach_cube_env.py:277) with the current ctrl, or with
 * n == 0 run mj_forward only (reach_cube_env.py:186,309).  Used by parity tests of the one-substep map. */
int lcr_substeps(LcrSim* sim, int n, void* stream);

/* Batched equivalent of the IK helper (inverse_kinematics, reach_cube_env.py:148-221; same text in all
 * envs; legacy interface SimulatedRobot.inverse_kinematics_reg, simulated_robot.py:139-186):
 * damped-least-squares iterations from the env's current arm qpos towards d_ee_target [n][3] f32,
 * result in d_q_out [n][6] f32.  Does not modify the simulation state (unlike the in-step IK). */
int lcr_ik(LcrSim* sim, const float* d_ee_target, float* d_q_out, void* stream);

/* Diagnostics of the last lcr_step/lcr_substeps: DEVICE array [n][LCR_NDIAG] int32 =
 * ncon, nefc, solver iterations (last substep), max nefc over the step, contacts dropped past the BIG caps of
 * lcr_model.h in the last substep (0 on every tested workload: the fast caps only route an env to the big workspace,
 * they never change a result), nan resets (mj_checkPos / mj_checkAcc -> mj_resetData). */
#define LCR_NDIAG 6
int lcr_get_diag(LcrSim* sim, int32_t* d_diag, void* stream);

/* Test hook: run mj_forward on the current state WITHOUT writing it back and dump the contact list:
 * d_contacts [n][LCR_MAXCON_BIG][12] float64 = pos[3], normal[3], dist, body1, body2, dim, mu, first row;
 * d_ncon [n] int32.  Lets the parity tests compare collision geometry with the oracle directly. */
int lcr_debug_contacts(LcrSim* sim, double* d_contacts, int32_t* d_ncon, void* stream);

/* Batched trajectory recorder: replaces HDF5_Recorder.capture_frame and the episode cut of RecordHDF5Wrapper.step
 * (gym_lowcostrobot/envs/wrappers/record_hdf5.py:40-45,116-137) for a whole batch, on the device.  Call once per step with the
 * step's d_obs [n][obs_dim] (columns 0:6 = arm_qpos, 6:12 = arm_qvel in every task), the d_actions [n][action_dim] that
 * produced it and the two done flags.  Appends the row qpos | qvel | action to the env's open trajectory
 * d_traj [n][horizon][12 + action_dim] (d_len [n] = rows so far, zero-initialised by the caller); an env whose episode ended
 * moves its trajectory to slot atomicAdd(d_count[0]) of the finished-episode pool d_pool [pool_cap][horizon][12 + action_dim],
 * d_pool_meta [pool_cap][2] = env, length (episodes beyond pool_cap are counted in d_count[1] and dropped) and starts a new
 * one.  The host drains the pool whenever it likes (read d_count, copy the slots, zero d_count[0]) -- no per-step sync.
 * The datasets the wrapper writes (observations/qpos, observations/qvel, action: record_hdf5.py:52-61) are column slices of
 * a pool slot.  Stateless: needs no simulator handle. */
int lcr_record_append(const float* d_obs, int obs_dim, const float* d_actions, int action_dim, const uint8_t* d_terminated,
                      const uint8_t* d_truncated, int n_envs, int horizon, float* d_traj, int32_t* d_len, float* d_pool,
                      int32_t* d_pool_meta, int32_t* d_count, int pool_cap, void* stream);

/* Image observations (reach_cube_env.py:109-112,288-292 and the same lines of the other envs: mujoco.Renderer +
 * update_scene(camera="camera_front" | "camera_top") + render(), 240 x 320 x 3 uint8) in two calls:
 *  lcr_body_poses  mj_kinematics on the current state of every env (nothing written back): d_poses [n][lcr_pose_slots()][12]
 *                  float32 = xpos[3] | xmat[9] (row-major) of the 7 arm bodies, then the boxes (cubes, PushCubeLoop's rails);
 *  lcr_render      stateless ray caster over convex geometry: d_geoms [n_geoms][LCR_RENDER_GEOM_WORDS] float32 = kind (0 hull,
 *                  1 box), pose slot, first half-space, half-space count, bounding-sphere centre[3] (body frame), radius,
 *                  half sizes[3] (the box itself, or the hull's bounding box about the same centre), rgb[3], 2 pad; d_planes [P][4] = half-spaces n . x + d <= 0 of the hulls in their body
 *                  frame; h_cameras (HOST) [n_cams][13] = position[3], rotation[9] (row-major, columns = camera x / y / z axes;
 *                  the camera looks along -z like MuJoCo's), fovy in degrees; d_images [n][n_cams][height][width][3] uint8.
 * The floor plane (0.1 m checker), the headlight and the scene's point light are built in; see csrc/lcr_render.cu for what
 * the images do and do not reproduce of MuJoCo's OpenGL renderer.  Returns 0 on success. */
#define LCR_RENDER_GEOM_WORDS 16
#define LCR_RENDER_MAXCAM 4
int lcr_pose_slots(const LcrSim* sim);
int lcr_body_poses(LcrSim* sim, float* d_poses, void* stream);
int lcr_render(const float* d_poses, int n_envs, int n_slots, const float* d_geoms, int n_geoms, const float* d_planes,
               const float* h_cameras, int n_cams, int height, int width, uint8_t* d_images, void* stream);

/* Debug hook (lockstep mode): from the next lcr_step on, every env writes the SM clock cycles it spent in each phase
 * of the step to d_clocks [n][10] int64 = begin/end, wait top, dynamics+broadphase, wait, narrowphase jobs, wait,
 * constraint rows, wait, Newton solve, integrate (sums over the substeps).  NULL switches it off again. */
int lcr_debug_phase_clocks(LcrSim* sim, long long* d_clocks);
/* Debug hook (flow mode): from the next lcr_step on, the warps of the flow kernel add the SM clock cycles they spent per
 * phase to d_stats [8] uint64 = BEGIN, DYN, JOB, COL, SOL, END, BIG, idle (never reset by the library).  NULL = off. */
int lcr_debug_flow_stats(LcrSim* sim, unsigned long long* d_stats);
/* Flow mode health check (synchronises the device): h_status[8]; [0] = envs that did not finish the last step (0 when the
 * step completed), [1] = watchdog code (0 = none; non-zero: a warp gave up waiting, the state is undefined), [2..7] =
 * debug words of the first failure (phase, env, ...). */
int lcr_flow_status(LcrSim* sim, int32_t* h_status);

int lcr_n_envs(const LcrSim* sim);
int lcr_kernel_launches(const LcrSim* sim); /* kernels launched by this handle so far */
const char* lcr_last_error(void);
const char* lcr_version(void);
/* sizeof(LcrModel) / sizeof(LcrEnvCfg) as compiled, so that FFI mirrors can verify their layout */
int lcr_sizeof_model(void);
int lcr_sizeof_cfg(void);

#ifdef __cplusplus
}
#endif
#endif /* LCRSIM_H_ */
