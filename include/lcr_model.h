/* lcr_model.h -- flat model constants shared by the C-ABI library, the CPU oracle and the
 * Python host (which mirrors this struct with ctypes in gym_lowcostrobot_b200/model.py).
 *
 * This is the compiled form of what the reference obtains from
 * mujoco.MjModel.from_xml_path(<scene>.xml)  (reference gym_lowcostrobot/envs/reach_cube_env.py:89,
 * push_cube_env.py:92, lift_cube_env.py:90, pick_place_cube_env.py:93, stack_two_cubes_env.py:90)
 * specialised to the one model family in scope: follower.xml (base_link + 6 hinged links,
 * 20 convex mesh geoms, 6 position servos, one ee site) + floor plane + 1 or 2 free cubes
 * (+ the four static box rails and the two goal regions of push_cube_loop.xml:40-48, read by
 * push_cube_loop_env.py:89,127-136).
 * It is a data format, not an algorithm: every field is a number read from the MJCF/STL
 * files or derived from them at qpos0 (invweight0, meaninertia).
 *
 * Indexing conventions
 *   arm body a = 0..6   : 0 = base_link (welded to the world), a = 1..6 = link_a, parent a-1
 *   joint / arm dof j   : 0..5, lives on arm body j+1
 *   cube c = 0..ncube-1 : free bodies; "body id" LCR_NABODY + c in contact records
 *   qpos = [arm q[6], cube0 pos[3] quat[4] (w,x,y,z), cube1 ...]         nq = 6 + 7*ncube
 *   qvel = [arm qd[6], cube0 lin[3] (world) ang[3] (body-local), ...]    nv = 6 + 6*ncube
 *   geom g              : 0..nmesh-1 arm meshes, nmesh = floor, nmesh+1+c = cube c,
 *                         nmesh+1+ncube+w = static wall box w (world body, axis aligned)
 *   box index           : cube c -> c, wall w -> ncube + w (convex-pair keys, contact parameter tables)
 */
#ifndef LCR_MODEL_H_
#define LCR_MODEL_H_

#include <stdint.h>

#define LCR_NARM 6
#define LCR_NABODY 7
#define LCR_MAXCUBE 2
#define LCR_MAXMESH 24
#define LCR_MAXWALL 4
#define LCR_MAXBOX (LCR_MAXCUBE + LCR_MAXWALL)
/* rows of the geom parameter table: meshes, floor, cubes, walls (the one scene with walls has one cube and 20 meshes,
 * so the table keeps its size: 20 + 1 + 1 + 4 <= 27) */
#define LCR_MAXGEOM (LCR_MAXMESH + 1 + LCR_MAXCUBE)
#define LCR_MAXPAIR 160
#define LCR_MAXNV (LCR_NARM + 6 * LCR_MAXCUBE)
#define LCR_MAXNQ (LCR_NARM + 7 * LCR_MAXCUBE)
/* Per-env caps of the contact list and of the constraint rows.  MuJoCo has no such caps (its arena is dynamic), so
 * the library keeps two workspace sizes: the FAST path (LCR_MAXCON / LCR_MAXEFC: what fits 16 envs into the shared
 * memory of one SM) and the BIG path (LCR_MAXCON_BIG / LCR_MAXEFC_BIG).  An env whose step would exceed a fast cap is
 * not written back by the fast kernels; it is queued and the same step is redone from the same start state by the
 * big-workspace kernel, so the caps never change a result.  Only what exceeds the BIG caps is dropped (in generation
 * order: limits, floor-cube, cube-cube, wall-cube, cube-mesh, wall-mesh, floor-mesh, mesh-mesh) and counted in
 * diag.overflow -- the tests assert that this never happens on the BASELINE workloads.  The oracle uses the BIG caps. */
#define LCR_MAXCON 32
#define LCR_MAXEFC 96
#define LCR_MAXCON_BIG 128
#define LCR_MAXEFC_BIG 384
/* entries of the per-env separating-axis cache of the convex narrowphase (performance only) */
#define LCR_NSA 16

enum { LCR_TASK_REACH = 0, LCR_TASK_PUSH = 1, LCR_TASK_LIFT = 2, LCR_TASK_PICK_PLACE = 3, LCR_TASK_STACK = 4, LCR_TASK_PUSH_LOOP = 5 };

typedef struct LcrModel {
  int32_t task, ncube, nq, nv;
  int32_t nmesh, nvert, npair, site_body;
  int32_t iterations, ls_iterations, pad0, pad1;
  double timestep, impratio, tolerance, ls_tolerance, meaninertia;
  double gravity[3];
  /* arm tree */
  double body_pos[LCR_NABODY][3], body_quat[LCR_NABODY][4];
  double body_ipos[LCR_NABODY][3], body_iquat[LCR_NABODY][4];
  double body_mass[LCR_NABODY], body_inertia[LCR_NABODY][3];
  double body_invweight0[LCR_NABODY][2];
  double jnt_axis[LCR_NARM][3], jnt_range[LCR_NARM][2];
  double jnt_armature[LCR_NARM], jnt_damping[LCR_NARM], jnt_frcrange[LCR_NARM][2];
  double jnt_solref[LCR_NARM][2], jnt_solimp[LCR_NARM][5], dof_invweight0[LCR_NARM];
  double act_kp[LCR_NARM], act_kv[LCR_NARM], act_ctrlrange[LCR_NARM][2];
  double site_pos[3];
  /* cubes */
  double cube_mass[LCR_MAXCUBE], cube_inertia[LCR_MAXCUBE][3], cube_size[LCR_MAXCUBE][3];
  double cube_invweight0[LCR_MAXCUBE][2];
  /* geoms */
  int32_t geom_condim[LCR_MAXGEOM], geom_priority[LCR_MAXGEOM];
  double geom_friction[LCR_MAXGEOM][3], geom_solref[LCR_MAXGEOM][2], geom_solimp[LCR_MAXGEOM][5];
  double geom_solmix[LCR_MAXGEOM];
  /* arm mesh geoms: convex hull vertex ranges in the pool passed beside the model (body frame),
   * body-frame bounding box (centre, half extents) and bounding-sphere radius about that centre */
  int32_t mesh_body[LCR_MAXMESH], mesh_vertadr[LCR_MAXMESH], mesh_vertnum[LCR_MAXMESH];
  double mesh_center[LCR_MAXMESH][3], mesh_half[LCR_MAXMESH][3], mesh_rbound[LCR_MAXMESH];
  /* geom centre MuJoCo gives a mesh geom = volume centroid of the mesh (body frame); interior point of MPR */
  double mesh_com[LCR_MAXMESH][3];
  /* candidate arm self-collision mesh pairs after the body-pair filter */
  int32_t pair_g1[LCR_MAXPAIR], pair_g2[LCR_MAXPAIR];
  /* static world boxes (push_cube_loop.xml:45-48: the rails that contain the cube): centre and half sizes, world frame,
   * axis aligned.  They collide with the cube (box-box) and with the meshes of the moving arm bodies (box-mesh). */
  int32_t nwall, pad2;
  double wall_pos[LCR_MAXWALL][3], wall_size[LCR_MAXWALL][3];
  /* goal regions of PushCubeLoop (push_cube_loop.xml:40-43; non-colliding boxes): geom_pos of goal_region_1 / _2 and
   * geom_size of goal_region_1, as read by push_cube_loop_env.py:127-133 */
  double goal_center[2][3], goal_size[3];
  /* body pos of the cubes in the MJCF = their qpos0 (what mj_resetData restores after a bad qacc) */
  double cube_pos0[LCR_MAXCUBE][3];
} LcrModel;

/* Per-env-class configuration = the reference Env constructor kwargs
 * (reference reach_cube_env.py:77-87, push_cube_env.py:79-90, lift_cube_env.py:77-88,
 * pick_place_cube_env.py:79-91, stack_two_cubes_env.py:78-88) plus TimeLimit(50)
 * (reference gym_lowcostrobot/__init__.py:9-37). */
typedef struct LcrEnvCfg {
  int32_t action_mode;      /* 0 = "joint", 1 = "ee" */
  int32_t block_gripper;    /* reference default: True for Reach/Push, False for Lift/PickPlace/Stack */
  int32_t reward_type;      /* 0 = "sparse", 1 = "dense" */
  int32_t n_substeps;       /* 20 */
  int32_t max_episode_steps;/* 50; <= 0 disables truncation */
  int32_t autoreset;        /* 0 = never (caller resets), 1 = next-step autoreset of done envs */
  int32_t collision_mask;   /* bit0 floor-cube, bit1 floor-mesh, bit2 cube-mesh, bit3 cube-cube, bit4 mesh-mesh, bit5 wall-cube, bit6 wall-mesh */
  int32_t exec_mode;        /* 0 = one fused kernel per step (one warp per CTA), 1 = phased (one small kernel per mj_step phase),
                             * 2 = lockstep (one kernel per step, CTAs of several envs aligned at the phase boundaries),
                             * 3 = flow (one persistent kernel per step, the phases of every env run from device-side queues) */
  double distance_threshold;/* 0.05 */
  double height_threshold;  /* 0.1 (Lift) */
  double cube_low[3], cube_high[3];     /* reset sampling box of the cube(s) */
  double target_low[3], target_high[3]; /* reset sampling box of the target (Push / PickPlace) */
} LcrEnvCfg;

#define LCR_COLLIDE_FLOOR_CUBE 1
#define LCR_COLLIDE_FLOOR_MESH 2
#define LCR_COLLIDE_CUBE_MESH 4
#define LCR_COLLIDE_CUBE_CUBE 8
#define LCR_COLLIDE_MESH_MESH 16
#define LCR_COLLIDE_WALL_CUBE 32
#define LCR_COLLIDE_WALL_MESH 64
#define LCR_COLLIDE_ALL 127

#endif /* LCR_MODEL_H_ */
