// lcr_flow.cuh -- "flow" execution of env.step(): ONE persistent kernel per step in which every env walks the phases of
// its 20 mj_step calls through device-side work queues, with the per-env workspace parked in HBM / L2 between phases
// and moved by the TMA unit (cp.async.bulk + mbarrier) instead of through registers.
//
// Why: the step is a long, data-dependent chain per env (Newton iterations, MPR runs).  The lockstep kernel keeps the
// warps of a CTA on the same code (the instruction cache needs that: a substep is ~110 KB of SASS) by CTA barriers and
// loses 60 % of its warp time waiting for the slowest env of the CTA; the phased chain gets the same code locality from
// one launch per phase and loses the tail of every one of its 164 launches.  Here the code locality comes from the SM:
// all warps of a CTA (one CTA per SM) serve the queue of the CTA's current phase (`role`), and a warp that finds that
// queue empty moves the CTA to the fullest one.  No env ever waits for another env, there is no barrier after start-up
// and no launch boundary inside the step:
//
//   k_sched_flow : ranks the envs by the constraint rows of their previous step, fills the BEGIN queue heaviest first
//   k_flow       : BEGIN (action map / IK / autoreset) -> 20 x [ DYN (integrate the previous substep, kinematics,
//                  inertia, bias, smooth forces, broadphase) -> JOB (one convex narrowphase candidate per item, any
//                  warp of the GPU) -> COL (contacts, constraint rows) -> SOL (Newton) ] -> END (integrate, reward,
//                  outputs incl. the packed record, state write-back)
//
// Envs that outgrow the fast workspace (LCR_MAXCON / LCR_MAXEFC) MIGRATE: the COL phase that hits the cap pushes the env
// (with the index of the substep it was about to build) to the BIG queue, and one of the few BIG CTAs of the grid
// resumes it from the parked substep-start state over the big workspace and finishes its step there (fused, one warp).
// Envs whose previous step already needed the big workspace start there.  Per-env arithmetic is identical to the other
// execution modes (bitwise identical state and outputs).
//
// Memory model notes.  Workspaces are written with st.global.cg and read by ANOTHER SM later in the same launch with
// cp.async.bulk (TMA unit, completion on the warp's mbarrier): the producer's lanes fence (generic -> async proxy, gpu
// scope) and lane 0 publishes the item with a release store; the consumer pops with relaxed loads and reads the
// workspace only through the TMA unit or ld.global.cg (L2) -- never through L1, which may hold a stale copy.  There is
// deliberately no __threadfence() in this kernel, see the note at st_release64.
#pragma once
#include "lcr_device.cuh"

namespace lcr {

// ---------------------------------------------------------------- PTX: mbarrier, bulk copies, proxy fences
DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DI void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DI bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared, completion counted in bytes on the mbarrier; size and both addresses are multiples of 16
DI void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DI void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
DI void bulk_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
DI void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DI void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
DI void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Publishing is done with release stores / release atomics (MEMBAR.ALL.GPU + the strong access) and consuming with
// relaxed loads followed by address-dependent reads that bypass L1 (the TMA unit, ld.global.cg): the kernel contains no
// __threadfence() / acquire fence.  Those compile to MEMBAR + CCTL.IVALL, an invalidation of the SM's whole L1 -- which
// also holds the local memory (register spills, call frames, saved convergence barriers) of every resident warp; with
// three or more busy warps per SM the kernel then ran on corrupted spills (lanes of a warp losing each other, bisected
// on B200 with tools/san_flow.py).  Nothing here needs the invalidation: no phase reads parked data through L1.
DI void st_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DI unsigned atom_add_release(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// a warp that waits this many SM clocks for one event (an item anywhere, a queue slot, a bulk copy) gives up and raises
// the error flag, which ends the kernel: a bug must never hang the device
#define LCR_FLOW_WATCHDOG 2000000000LL

// ---------------------------------------------------------------- work queues
// (queue ids, FlowQ: lcr_device.cuh)  item = env (20 bits) | aux (10 bits) << 20 | express << 30
DI unsigned fq_item(int env, int aux, int express) { return (unsigned)env | ((unsigned)aux << 20) | ((unsigned)express << 30); }
DI int fq_env(unsigned it) { return (int)(it & 0xfffffu); }
DI int fq_aux(unsigned it) { return (int)((it >> 20) & 0x3ffu); }
DI int fq_express(unsigned it) { return (int)((it >> 30) & 1u); }

DI unsigned* fq_head(const FlowQ& q, int i) { return q.ctl + 64 * i; }
DI unsigned* fq_tail(const FlowQ& q, int i) { return q.ctl + 64 * i + 32; }
DI unsigned* fq_remaining(const FlowQ& q) { return q.ctl + 64 * LCR_FQ_NQ; }       // envs that have not finished this step
DI unsigned* fq_error(const FlowQ& q) { return q.ctl + 64 * LCR_FQ_NQ + 32; }      // watchdog: a warp gave up waiting
// relaxed gpu-scope accesses of the queue words (C++ volatile would be system scope: LD.E.STRONG.SYS)
DI unsigned ldv(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Bounded multi-producer / multi-consumer rings with a sequence number per slot (Vyukov): slot = seq << 32 | item, one
// 64-bit word that is read and written whole.  Ticket t owns slot t & mask when the slot's seq == t (free for this lap);
// the producer publishes (t + 1, item), the consumer of ticket t takes it and frees the slot for the next lap with
// seq = t + cap.  head / tail are never reset (32-bit modular arithmetic), so a slow consumer or producer of an
// earlier lap can never be overtaken -- it only makes the owner of the next lap wait.
DI unsigned long long ldv64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DI void stv64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
DI unsigned long long* fq_slot(const FlowQ& q, int qi, unsigned ticket) {
  return reinterpret_cast<unsigned long long*>(q.ring[qi]) + (ticket & q.mask[qi]);
}
// one lane; the caller has fenced its data.  Ticket already taken (t): publish the item.
DI void fq_publish(const FlowQ& q, int qi, unsigned t, unsigned item) {
  unsigned long long* slot = fq_slot(q, qi, t);
  long long t0 = 0;
  while ((unsigned)(ldv64(slot) >> 32) != t) {  // the consumer of the previous lap has not freed the slot yet (ring nearly full)
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > LCR_FLOW_WATCHDOG) { atomicExch(fq_error(q), 4u); return; }
  }
  st_release64(slot, ((unsigned long long)(t + 1u) << 32) | item);  // release: the caller's data before the item
}
DI void fq_push(const FlowQ& q, int qi, unsigned item) { fq_publish(q, qi, atomicAdd(fq_tail(q, qi), 1u), item); }
// one lane; LCR_FQ_EMPTY if the queue has nothing for us
DI unsigned fq_pop(const FlowQ& q, int qi) {
  for (;;) {
    const unsigned h = ldv(fq_head(q, qi)), t = ldv(fq_tail(q, qi));
    if ((int)(t - h) <= 0) return LCR_FQ_EMPTY;
    if (atomicCAS(fq_head(q, qi), h, h + 1) != h) continue;
    unsigned long long* slot = fq_slot(q, qi, h);
    unsigned long long v;
    long long t0 = 0;
    // the producer of ticket h is between its atomicAdd and its store
    while ((unsigned)((v = ldv64(slot)) >> 32) != h + 1u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > LCR_FLOW_WATCHDOG) { atomicExch(fq_error(q), 2u); return LCR_FQ_EMPTY; }
    }
    stv64(slot, (unsigned long long)(h + q.mask[qi] + 1u) << 32);  // free for the next lap
    return (unsigned)v;
  }
}
DI unsigned fq_pop_phase(const FlowQ& q, int phase) {
  const unsigned it = fq_pop(q, 2 * phase);
  return it != LCR_FQ_EMPTY ? it : fq_pop(q, 2 * phase + 1);
}
#ifdef LCR_FLOW_DEBUG
#define FLOW_COUNT(q, k) atomicAdd((q).ctl + 960 + (k), 1u)
#else
#define FLOW_COUNT(q, k) ((void)0)
#endif
DI void fq_push_phase(const FlowQ& q, int phase, int env, int aux, int express) {
  FLOW_COUNT(q, phase);
  fq_push(q, 2 * phase + (express ? 0 : 1), fq_item(env, aux, express));
}

// ---------------------------------------------------------------- staging through the TMA unit
struct Stage {
  unsigned long long* bar;  // this warp's mbarrier
  unsigned parity;
};
// regions of the parked workspace [off0, off1) -> the same offsets of the shared-memory workspace; all lanes call
template <typename WsT, int N>
DI void stage_in(Stage& sg, WsT& w, const WsT* g, const int (&off)[N][2], const FlowQ& fq) {
  fence_async_smem();  // earlier generic accesses of this shared memory are ordered before the async-proxy writes
  __syncwarp();
  if (LANE == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int k = 0; k < N; k++) tot += (unsigned)(off[k][1] - off[k][0]);
    mbar_expect_tx(sg.bar, tot);
#pragma unroll
    for (int k = 0; k < N; k++)
      if (off[k][1] > off[k][0])
        bulk_g2s(reinterpret_cast<unsigned char*>(&w) + off[k][0], reinterpret_cast<const unsigned char*>(g) + off[k][0],
                 (unsigned)(off[k][1] - off[k][0]), sg.bar);
  }
  long long t0 = 0;
  while (!mbar_try_wait(sg.bar, sg.parity)) {
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > LCR_FLOW_WATCHDOG) { if (LANE == 0) atomicExch(fq_error(fq), 3u); break; }
  }
  sg.parity ^= 1u;
}
// regions of the shared-memory workspace -> the parked one, performed GPU-wide on return; all lanes call.
// 128-bit st.global.cg from all lanes, not a bulk store: completing a bulk store (cp.async.bulk.wait_group) costs a
// CCTL.IVALL like a __threadfence() (see above).  `fence.proxy.async` afterwards is both the generic -> async proxy fence
// the consumer's TMA read needs from the writer and a gpu-scope MEMBAR without invalidation (SASS: MEMBAR.ALL.GPU +
// FENCE.VIEW.ASYNC), so when lane 0 publishes the item after the __syncwarp every lane's part has reached L2.
template <typename WsT, int N>
DI void stage_out(const Stage&, const WsT& w, WsT* g, const int (&off)[N][2]) {
  __syncwarp();
  asm volatile("" ::: "memory");
#pragma unroll
  for (int k = 0; k < N; k++) {
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(g) + off[k][0]);
    const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(&w) + off[k][0]);
    for (int i = LANE; i < (off[k][1] - off[k][0]) / 16; i += 32) __stcg(d + i, s[i]);
  }
  fence_async_all();
  __syncwarp();
}
// plain 128-bit copies that bypass L1 (BIG CTAs: the parked fast workspace -> a big one with different offsets)
DI void copy_cg(void* dst, const void* src, int bytes) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = LANE; i < bytes / 16; i += 32) d[i] = __ldcg(s + i);
}

#define LCR_FOFF(member) ((int)offsetof(WsT, member))
// The phase handlers are real functions, not inlined into the dispatch loop of k_flow: one body of ~100 KB with every
// phase inlined into a single switch both spilled three times as much (364 vs 136 bytes) and, built with nvcc 12.9,
// produced corrupt workspaces as soon as three or more warps of a CTA were active (bisected on B200 with
// tools/san_flow.py; the same source with these calls out of line is bit-identical to the fused kernel).
#define LCR_FLOW_FN __device__ __noinline__

// ---------------------------------------------------------------- scheduler
// One CTA.  Zeroes the queue control block, ranks the envs by the constraint rows of their previous step (diag[3];
// contacts persist, so it predicts the cost of this one; envs that will only be auto-reset are the cheapest) and fills
// the BEGIN queue heaviest first.  Envs with >= t_big rows start on the BIG path, envs with >= t_hi rows travel express.
template <typename T>
__global__ void __launch_bounds__(1024) k_sched_flow(DevState<T> s, FlowQ fq, int t_hi, int t_big) {
  __shared__ int hist[LCR_NBUCKET], start[LCR_NBUCKET];
  __shared__ unsigned base_big, base_begin;
  // (all queues are empty between steps: head == tail, every slot free for its next lap; the counters run on)
  if (threadIdx.x < 128) fq.ctl[64 * LCR_FQ_NQ + threadIdx.x] = 0;  // remaining | error | debug words | debug counters
  if (threadIdx.x < LCR_NBUCKET) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < s.n; e += blockDim.x) {
    const int32_t* ib = s.ib + (size_t)e * LCR_IB_WORDS;
    const int prev = ib[LCR_NINT + 3];
    int key = ib[1] ? 0 : 1 + prev / 8;
    key = key < LCR_NBUCKET - 1 ? key : LCR_NBUCKET - 2;
    if (!ib[1] && prev >= t_big) key = LCR_NBUCKET - 1;
    atomicAdd(&hist[key], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = LCR_NBUCKET - 1; k >= 0; k--) { start[k] = acc; acc += hist[k]; }
    const int nbig = hist[LCR_NBUCKET - 1];
    base_big = *fq_tail(fq, 2 * FQ_BIG);
    base_begin = *fq_tail(fq, 2 * FQ_BEGIN + 1);
    *fq_tail(fq, 2 * FQ_BIG) = base_big + (unsigned)nbig;
    *fq_tail(fq, 2 * FQ_BEGIN + 1) = base_begin + (unsigned)(s.n - nbig);
    *fq_remaining(fq) = (unsigned)s.n;
  }
  __syncthreads();
  const int nbig = hist[LCR_NBUCKET - 1];
  for (int e = threadIdx.x; e < s.n; e += blockDim.x) {
    const int32_t* ib = s.ib + (size_t)e * LCR_IB_WORDS;
    const int prev = ib[LCR_NINT + 3];
    int key = ib[1] ? 0 : 1 + prev / 8;
    key = key < LCR_NBUCKET - 1 ? key : LCR_NBUCKET - 2;
    const bool big = !ib[1] && prev >= t_big;
    if (big) key = LCR_NBUCKET - 1;
    const int r = atomicAdd(&start[key], 1);
    const int express = (!ib[1] && prev >= t_hi) ? 1 : 0;
    if (big) fq_publish(fq, 2 * FQ_BIG, base_big + (unsigned)r, fq_item(e, 0, 1));
    else fq_publish(fq, 2 * FQ_BEGIN + 1, base_begin + (unsigned)(r - nbig), fq_item(e, 0, express));
  }
}

// ---------------------------------------------------------------- the phases (one warp, workspace slot `w` in shared memory)
#define LCR_FLOW_FUSE_DYN_COL 1  // an env without convex candidates builds its constraint rows in the DYN warp (one hop less)
#define LCR_FLOW_FUSE_COL_SOL 2  // the warp that built the rows also solves (one hop less, larger code per SM)
#define LCR_FLOW_DUAL_SCAN 4     // JOB phase: scan both hulls of a mesh pair with their loads in flight together

template <typename T, int S>
struct FlowCtx {
  const DevModel<T>& m;
  const T* __restrict__ verts;
  DevState<T> s;
  Ws<T, S>* gws;
  const StepIO& io;
  const FlowQ& fq;
  int flags;
};

// debug: every parked workspace carries the id of its env (w.skip = env + 1, written by BEGIN); a phase that staged in
// something else read a workspace that was not (yet) written -- record it and stop the kernel
template <typename T, int S> DI bool flow_tag_ok(const FlowCtx<T, S>& c, const Ws<T, S>& w, int env, int phase) {
  if (w.skip == env + 1) return true;
  if (LANE == 0 && atomicCAS(fq_error(c.fq), 0u, 7u) == 0u) {
    unsigned* d = fq_error(c.fq) + 1;
    d[0] = (unsigned)phase; d[1] = (unsigned)env; d[2] = (unsigned)w.skip; d[3] = (unsigned)w.substep; d[4] = blockIdx.x; d[5] = threadIdx.x >> 5;
  }
  return false;
}
#ifdef LCR_FLOW_DEBUG
// debug: fields that steer control flow must be identical in every lane and within their ranges
template <typename T, int S> DI void flow_check(const Ws<T, S>& w, int env, int tag) {
  const int a = w.substep, b = w.redo_forward, c = w.nefc, d = w.ncon, e = w.skip, f = w.ovf;
  const bool bad = a != __shfl_sync(FULLMASK, a, 0) || b != __shfl_sync(FULLMASK, b, 0) || c != __shfl_sync(FULLMASK, c, 0) ||
                   d != __shfl_sync(FULLMASK, d, 0) || e != env + 1 || a < 0 || a > 20 || (unsigned)b > 1u || (unsigned)c > 96u || (unsigned)d > 32u || (unsigned)f > 1u;
  if (__any_sync(FULLMASK, bad)) {
    if (bad) printf("flow_check tag %d env %d blk %d warp %d lane %d: substep %d redo %d nefc %d ncon %d skip %d ovf %d\n", tag, env, blockIdx.x, threadIdx.x >> 5, LANE, a, b, c, d, e, f);
    __trap();
  }
}
#define FLOW_CHECK(w, env, tag) flow_check(w, env, tag)
#else
#define FLOW_CHECK(w, env, tag) ((void)0)
#endif
template <typename T, int S> DI void flow_done(const FlowCtx<T, S>& c) {
  if (LANE == 0) atom_add_release(fq_remaining(c.fq), 0xffffffffu);
}

// constraint rows are built; ship them to a solver warp (or hand the env to the BIG path if a cap was hit)
template <typename T, int S>
LCR_FLOW_FN void flow_after_rows(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express, bool state_parked, bool dyn_in_smem);

// cache_in_smem: this warp also holds (and may have changed) the counts + separating-axis cache block of the env
template <typename T, int S>
LCR_FLOW_FN void flow_solve(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express, bool cache_in_smem) {
  typedef Ws<T, S> WsT;
  const DevModel<T>& m = c.m;
  solve_constraints<T, S, false>(w, m, solver_tol<T>(m));
  if (check_acc(w, m)) { if (LANE == 0) w.redo_forward = 1; }
  int k = 0;
  if (LANE == 0) { k = w.substep + 1; w.substep = k; }
  k = __shfl_sync(FULLMASK, k, 0);
  __syncwarp();
  FLOW_CHECK(w, env, 41);
  // writes: state (warm start, diag; qpos / qvel if the env was reset), dynamics vectors (qacc), flags
  const int out[3][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(M), LCR_FOFF(H)}, {cache_in_smem ? LCR_FOFF(ncon) : LCR_FOFF(cand_key), LCR_FOFF(J)}};
  stage_out(sg, w, c.gws + env, out);
  if (LANE == 0) fq_push_phase(c.fq, k >= m.n_substeps ? FQ_END : FQ_DYN, env, 0, express);
}

template <typename T, int S>
LCR_FLOW_FN void flow_after_rows(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express, bool state_parked, bool dyn_in_smem) {
  typedef Ws<T, S> WsT;
  if (w.ovf) {
    // the env needs the big workspace: it resumes there from the start of this substep (state block of the parked workspace)
    if (!state_parked) {
      const int out[1][2] = {{0, LCR_FOFF(xpos)}};
      stage_out(sg, w, c.gws + env, out);
    }
    if (LANE == 0) fq_push_phase(c.fq, FQ_BIG, env, w.substep + 1, 1);
    return;
  }
  const int nefc = w.nefc;
  FLOW_CHECK(w, env, dyn_in_smem ? 12 : 31);
  const int jend = LCR_FOFF(J) + ((nefc * WsT::JS * (int)sizeof(T) + 15) / 16) * 16;
  if ((c.flags & LCR_FLOW_FUSE_COL_SOL) && dyn_in_smem) {
    // everything the solver reads is already in this warp's shared memory
    flow_solve(c, sg, w, env, express, true);
    return;
  }
  if (dyn_in_smem) {
    // DYN + COL in one warp: state, kinematics, dynamics vectors | contacts, row parameters | row maps, counts, cache, candidates | J
    const int out[4][2] = {{0, LCR_FOFF(H)}, {LCR_FOFF(c_pos), LCR_FOFF(e_jar)}, {LCR_FOFF(e_unit), LCR_FOFF(J)}, {LCR_FOFF(J), jend}};
    stage_out(sg, w, c.gws + env, out);
  } else {
    // writes: diag (state block), contacts, row parameters, row -> contact maps, counts + cache, J
    const int out[4][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(c_pos), LCR_FOFF(e_jar)}, {LCR_FOFF(e_unit), LCR_FOFF(cand_key)}, {LCR_FOFF(J), jend}};
    stage_out(sg, w, c.gws + env, out);
  }
  if (LANE == 0) fq_push_phase(c.fq, FQ_SOL, env, nefc, express);
  (void)sg;
}

template <typename T, int S>
LCR_FLOW_FN void flow_begin(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express) {
  typedef Ws<T, S> WsT;
  const DevModel<T>& m = c.m;
  load_state(w, c.s, env);
  const bool go = env_step_begin(w, m, c.verts, c.io, env);
  if (w.ovf) {  // the IK / reset forward passes outgrew the fast workspace: redo the whole step on the BIG path (state untouched)
    if (LANE == 0) fq_push_phase(c.fq, FQ_BIG, env, 0, 1);
    return;
  }
  if (!go) {  // auto-reset instead of a step: done
    store_state(w, c.s, env);
    flow_done(c);
    return;
  }
  if (LANE == 0) { w.skip = env + 1; w.redo_forward = 0; w.nefc = 0; w.ncon = 0; w.nlim = 0; w.substep = 0; w.jobs_left = 0; }
  __syncwarp();
  FLOW_CHECK(w, env, 1);
  const int out[2][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(ncon), LCR_FOFF(J)}};
  stage_out(sg, w, c.gws + env, out);
  if (LANE == 0) fq_push_phase(c.fq, FQ_DYN, env, 0, express);
  (void)sg;
}

// [integrate the previous substep] -> checks -> kinematics -> inertia / bias -> smooth forces -> broadphase
template <typename T, int S>
LCR_FLOW_FN void flow_dyn(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express) {
  typedef Ws<T, S> WsT;
  const DevModel<T>& m = c.m;
  {
    // reads: state, dynamics vectors (M, qacc for the integration of the previous substep), counts + cache, candidate block
    const int in[3][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(M), LCR_FOFF(H)}, {LCR_FOFF(ncon), LCR_FOFF(J)}};
    stage_in(sg, w, c.gws + env, in, c.fq);
  }
  if (!flow_tag_ok(c, w, env, FQ_DYN)) return;
  FLOW_CHECK(w, env, 10);
  bool redone = false;
  if (w.substep > 0) {
    if (w.redo_forward) { forward(w, m, c.verts); if (LANE == 0) { w.redo_forward = 0; w.ovf = 0; } __syncwarp(); redone = true; }
    integrate(w, m);
  }
  check_state(w, m);
  kinematics(w, m);
  inertia_and_bias(w, m);
  smooth_forces(w, m);
  collect_candidates(w, m);
  __syncwarp();
  const int nc = w.ncand < WsT::MAXCAND ? w.ncand : WsT::MAXCAND;
  if (nc == 0 && (c.flags & LCR_FLOW_FUSE_DYN_COL)) {
    make_constraints(w, m, c.verts, true);
    __syncwarp();
    flow_after_rows(c, sg, w, env, express, false, true);
    return;
  }
  if (LANE == 0) w.jobs_left = nc;
  __syncwarp();
  FLOW_CHECK(w, env, 11);
  // writes: state, kinematics, dynamics vectors, candidate block (+ the cache if mj_forward was re-run)
  const int out[2][2] = {{0, LCR_FOFF(H)}, {redone ? LCR_FOFF(ncon) : LCR_FOFF(cand_key), LCR_FOFF(J)}};
  stage_out(sg, w, c.gws + env, out);
  if (nc == 0) {
    if (LANE == 0) fq_push_phase(c.fq, FQ_COL, env, 0, express);
  } else {
    // one item per candidate: a single ticket range, the lanes fill it
    const int qi = 2 * FQ_JOB + (express ? 0 : 1);
    unsigned t0 = 0;
    if (LANE == 0) t0 = atomicAdd(fq_tail(c.fq, qi), (unsigned)nc);
    t0 = __shfl_sync(FULLMASK, t0, 0);
    for (int k = LANE; k < nc; k += 32) { FLOW_COUNT(c.fq, FQ_JOB); fq_publish(c.fq, qi, t0 + (unsigned)k, fq_item(env, k, express)); }
  }
}

// one convex candidate of one env; the last job of the env sends it on to COL
template <typename T, int S>
LCR_FLOW_FN void flow_job(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int k, int express) {
  typedef Ws<T, S> WsT;
  {
    // reads: kinematics (poses, bounding-sphere centres), counts + cache + candidate keys
    const int in[2][2] = {{LCR_FOFF(xpos), LCR_FOFF(M)}, {LCR_FOFF(ncon), LCR_FOFF(J)}};
    stage_in(sg, w, c.gws + env, in, c.fq);
  }
  if (!flow_tag_ok(c, w, env, FQ_JOB)) return;
  FLOW_CHECK(w, env, 20);
  T r[8];
  if (c.flags & LCR_FLOW_DUAL_SCAN) narrowphase_job<T, S, true>(w, c.m, c.verts, w.cand_key[k], r);
  else narrowphase_job<T, S, false>(w, c.m, c.verts, w.cand_key[k], r);
  WsT& g = c.gws[env];
  if (LANE == 0) {  // (the result is warp-uniform: one lane writes it, so that its release below covers all of it)
    float4* dst = reinterpret_cast<float4*>(cand_res(g)[k]);
    if (sizeof(T) == 4) {
      __stcg(dst, make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]));
      __stcg(dst + 1, make_float4((float)r[4], (float)r[5], (float)r[6], (float)r[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) cand_res(g)[k][j] = r[j];
    }
    fence_async_global();  // generic-proxy writes, read by the COL warp through the async proxy
    const unsigned left = atom_add_release(reinterpret_cast<unsigned*>(&g.jobs_left), 0xffffffffu);
    if (left == 1u) fq_push_phase(c.fq, FQ_COL, env, 0, express);
  }
}

template <typename T, int S>
LCR_FLOW_FN void flow_col(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int express) {
  typedef Ws<T, S> WsT;
  {
    // reads: state, kinematics, the job results (they alias e_w / e_g / e_p), counts + cache, candidate block
    const int in[3][2] = {{0, LCR_FOFF(M)}, {LCR_FOFF(e_w), LCR_FOFF(e_unit)}, {LCR_FOFF(ncon), LCR_FOFF(J)}};
    stage_in(sg, w, c.gws + env, in, c.fq);
  }
  if (!flow_tag_ok(c, w, env, FQ_COL)) return;
  FLOW_CHECK(w, env, 30);
  make_constraints(w, c.m, c.verts, true);
  __syncwarp();
  if (!w.ovf && (c.flags & LCR_FLOW_FUSE_COL_SOL)) {
    // the solver also reads the dynamics vectors
    const int in[1][2] = {{LCR_FOFF(M), LCR_FOFF(H)}};
    stage_in(sg, w, c.gws + env, in, c.fq);
    flow_solve(c, sg, w, env, express, true);  // (contacts / rows stay local; the cache changed by make_constraints is written back)
    return;
  }
  flow_after_rows(c, sg, w, env, express, true, false);
}

template <typename T, int S>
LCR_FLOW_FN void flow_sol(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env, int nefc, int express) {
  typedef Ws<T, S> WsT;
  {
    // reads: state (warm start), dynamics vectors, contact scalars (not positions / frames), row parameters, counts, flags, live rows of J
    const int jend = LCR_FOFF(J) + ((nefc * WsT::JS * (int)sizeof(T) + 15) / 16) * 16;
    const int in[6][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(M), LCR_FOFF(H)}, {LCR_FOFF(c_dist), LCR_FOFF(e_jar)},
                          {LCR_FOFF(ncon), LCR_FOFF(sa_dir)}, {LCR_FOFF(cand_key), LCR_FOFF(J)}, {LCR_FOFF(J), jend}};
    stage_in(sg, w, c.gws + env, in, c.fq);
  }
  if (!flow_tag_ok(c, w, env, FQ_SOL)) return;
  FLOW_CHECK(w, env, 40);
  flow_solve(c, sg, w, env, express, false);
}

template <typename T, int S>
LCR_FLOW_FN void flow_end(const FlowCtx<T, S>& c, Stage& sg, Ws<T, S>& w, int env) {
  typedef Ws<T, S> WsT;
  const DevModel<T>& m = c.m;
  {
    const int in[3][2] = {{0, LCR_FOFF(xpos)}, {LCR_FOFF(M), LCR_FOFF(H)}, {LCR_FOFF(ncon), LCR_FOFF(J)}};
    stage_in(sg, w, c.gws + env, in, c.fq);
  }
  if (!flow_tag_ok(c, w, env, FQ_END)) return;
  FLOW_CHECK(w, env, 50);
  if (w.redo_forward) forward(w, m, c.verts);
  integrate(w, m);
  env_step_end(w, m, c.io, env);
  store_state(w, c.s, env);
  flow_done(c);
}

// BIG path: the whole step (aux == 0) or the rest of it from substep aux - 1 (migrated env), fused, over the big workspace
template <typename T, int S>
LCR_FLOW_FN void flow_big(const FlowCtx<T, S>& c, Ws<T, S | LCR_NC_BIG>& wb, int env, int aux) {
  typedef Ws<T, S> WsT;
  typedef Ws<T, S | LCR_NC_BIG> WsB;
  const DevModel<T>& m = c.m;
  if (aux == 0) {
    load_state(wb, c.s, env);
    env_step(wb, m, c.verts, c.io, env);
  } else {
    // the state block and the separating-axis cache have the same layout in both workspaces
    static_assert(offsetof(WsT, xpos) == offsetof(WsB, xpos), "state block layout");
    copy_cg(&wb, c.gws + env, (int)offsetof(WsT, xpos));
    copy_cg(wb.sa_dir, reinterpret_cast<const unsigned char*>(c.gws + env) + offsetof(WsT, sa_dir), WsT::SA_BYTES);
    if (LANE == 0) { wb.ovf = 0; wb.ncand = 0; wb.redo_forward = 0; }
    __syncwarp();
#pragma unroll 1
    for (int k = aux - 1; k < m.n_substeps; k++) substep(wb, m, c.verts);
    env_step_end(wb, m, c.io, env);
  }
  store_state(wb, c.s, env);
  flow_done(c);
}

#define LCR_LQ 32  // entries of a CTA's local work ring (power of two)
// position of the n-th (0-based) set bit of m
DI int nth_bit(unsigned m, int n) {
  for (int i = 0; i < n; i++) m &= m - 1u;
  return __ffs((int)m) - 1;
}
template <typename U> DI volatile U& vol_ref(U& x) { return *reinterpret_cast<volatile U*>(&x); }

// ---------------------------------------------------------------- the kernel
// grid = one CTA per SM; the last `nbigcta` CTAs serve the BIG queue (as many one-warp slots as big workspaces fit their
// shared memory), the others run blockDim.x / 32 warps with one fast workspace slot each.
// stats (optional, [8] per phase: busy clocks; [7] idle clocks): debug hook.
template <typename T, int S>
__global__ void __launch_bounds__(512, 1) k_flow(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                                 Ws<T, S>* __restrict__ gws, StepIO io, FlowQ fq, int nbigcta, int nbigslot, int flags,
                                                 unsigned long long* __restrict__ stats) {
  __shared__ unsigned long long mbar[16];
  __shared__ int role, scout_lock, quit, n_idle, lq_head, lq_tail, lq_ph[LCR_LQ];
  __shared__ unsigned lq_item[LCR_LQ];
  const int warp = threadIdx.x >> 5, lane = LANE;
  const bool bigcta = (int)blockIdx.x >= (int)gridDim.x - nbigcta;
  FlowCtx<T, S> c{*dm, verts, s, gws, io, fq, flags};
  long long tb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = stats ? clock64() : 0;
#define LCR_FLOW_TICK(k) do { if (stats) { const long long t1_ = clock64(); tb[k] += t1_ - t0; t0 = t1_; } } while (0)
  if (bigcta) {
    if (warp >= nbigslot) return;
    Ws<T, S | LCR_NC_BIG>& wb = reinterpret_cast<Ws<T, S | LCR_NC_BIG>*>(lcr_smem)[warp];
    long long idle0 = 0;
    for (;;) {
      unsigned item = LCR_FQ_EMPTY;
      if (lane == 0) item = fq_pop_phase(fq, FQ_BIG);
      item = __shfl_sync(FULLMASK, item, 0);
      if (item == LCR_FQ_EMPTY) {
        int stop = 0;
        if (lane == 0) {
          stop = ldv(fq_remaining(fq)) == 0 || ldv(fq_error(fq)) != 0;
          const long long now = clock64();
          if (idle0 == 0) idle0 = now;
          else if (now - idle0 > LCR_FLOW_WATCHDOG) { atomicExch(fq_error(fq), 1u); stop = 1; }
        }
        if (__shfl_sync(FULLMASK, stop, 0)) break;
        __nanosleep(2000);
        continue;
      }
      idle0 = 0;
      LCR_FLOW_TICK(7);
      flow_big(c, wb, fq_env(item), fq_aux(item));
      LCR_FLOW_TICK(FQ_BIG);
    }
  } else {
    Ws<T, S>& w = reinterpret_cast<Ws<T, S>*>(lcr_smem)[warp];
    Stage sg{&mbar[warp], 0u};
    if (lane == 0) mbar_init(sg.bar, 1);
    if (threadIdx.x == 0) { role = FQ_BEGIN; scout_lock = 0; n_idle = 0; quit = 0; lq_head = 0; lq_tail = 0; }
    fence_async_smem();  // make the initialised barriers visible to the async proxy
    __syncthreads();
    // Work distribution.  Thousands of warps popping the same few queue words would serialise on their L2 lines (measured:
    // the step got SLOWER with every warp added), so only ONE warp per CTA at a time -- the scout, holder of scout_lock --
    // talks to the global queues: it sizes up the queues and claims a batch of items of one phase with a single CAS (at
    // least one per idle warp of the CTA, at most the CTA's fair share of that queue) into the CTA's LOCAL ring in shared
    // memory.  Every warp takes its work from the local ring; idle warps sleep on it without any global traffic.
    long long idle0 = 0;
    unsigned backoff = 128;
#define LCR_VOL(x) vol_ref(x)
    const int nwarps = (int)(blockDim.x >> 5), nnormal = (int)gridDim.x - nbigcta;
    for (;;) {
      // ---- local ring: read entry, then claim it (the entry cannot be rewritten before lq_head has passed it)
      int ph = -1;
      unsigned item = 0;
      if (lane == 0) {
        for (;;) {
          const int h = LCR_VOL(lq_head);
          if (h == LCR_VOL(lq_tail)) break;
          const unsigned it = LCR_VOL(lq_item[h & (LCR_LQ - 1)]);
          const int p2 = LCR_VOL(lq_ph[h & (LCR_LQ - 1)]);
          if (atomicCAS(&lq_head, h, h + 1) == h) { item = it; ph = p2; break; }
        }
      }
      ph = __shfl_sync(FULLMASK, ph, 0);
      item = __shfl_sync(FULLMASK, item, 0);
      if (ph >= 0) {
        idle0 = 0;
        backoff = 128;
        LCR_FLOW_TICK(7);
        const int env = fq_env(item), aux = fq_aux(item), ex = fq_express(item);
        if (lane == 0) FLOW_COUNT(fq, 16 + ph);
        switch (ph) {
          case FQ_BEGIN: flow_begin(c, sg, w, env, ex); break;
          case FQ_DYN: flow_dyn(c, sg, w, env, ex); break;
          case FQ_JOB: flow_job(c, sg, w, env, aux, ex); break;
          case FQ_COL: flow_col(c, sg, w, env, ex); break;
          case FQ_SOL: flow_sol(c, sg, w, env, aux, ex); break;
          default: flow_end(c, sg, w, env); break;
        }
        LCR_FLOW_TICK(ph);
        continue;
      }
      // ---- the local ring is empty: become the scout, or sleep on the ring (shared memory only)
      int scout = 0;
      if (lane == 0) {
        atomicAdd(&n_idle, 1);
        scout = atomicCAS(&scout_lock, 0, 1) == 0;
      }
      scout = __shfl_sync(FULLMASK, scout, 0);
      if (!scout) {
        int q = 0;
        if (lane == 0) {
          for (;;) {
            __nanosleep(100);
            q = LCR_VOL(quit);
            if (LCR_VOL(lq_head) != LCR_VOL(lq_tail) || q || LCR_VOL(scout_lock) == 0) break;
          }
          q = q && LCR_VOL(lq_head) == LCR_VOL(lq_tail);
          atomicSub(&n_idle, 1);
        }
        if (__shfl_sync(FULLMASK, q, 0)) break;
        continue;
      }
      // ---- scout
      int stop = 0;
      for (;;) {
        int want = __shfl_sync(FULLMASK, LCR_VOL(n_idle), 0);  // idle warps of the CTA (includes this one)
        want = want < 1 ? 1 : (want > nwarps ? nwarps : want);
        unsigned d = 0;
        if (lane < 2 * FQ_BIG) d = ldv(fq_tail(fq, lane)) - ldv(fq_head(fq, lane));
        if ((int)d < 0) d = 0;
        const unsigned dp = d + __shfl_xor_sync(FULLMASK, d, 1);  // items of the lane's phase (both priorities)
        // stay with the CTA's phase while it has work (instruction cache), else the fullest queue (ties: the later phase)
        const int cur = __shfl_sync(FULLMASK, LCR_VOL(role), 0);
        const unsigned dcur = __shfl_sync(FULLMASK, dp, 2 * cur);
        unsigned key = (lane < 2 * FQ_BIG && dp > 0) ? ((dp << 3) | (unsigned)(lane >> 1)) : 0u;
        key = __reduce_max_sync(FULLMASK, key);
        if (key != 0) {
          const int p = dcur > 0 ? cur : (int)(key & 7u);
          const int depth = (int)__shfl_sync(FULLMASK, dp, 2 * p);
          // batch: one item per idle warp, or the CTA's fair share of the queue if that is more; bounded by the free ring slots
          const int fill = __shfl_sync(FULLMASK, LCR_VOL(lq_tail) - LCR_VOL(lq_head), 0);
          int room = LCR_LQ - fill;
          // (deep queue: a few more than the idle warps need, so that warps finishing soon find work in the ring without
          //  waiting for a scout round; shallow queue -- the tail of the step --: no hoarding)
          int share = want;
          if (depth >= 4 * nnormal) {
            share = want + nwarps / 2;
            if (depth / nnormal > share) share = depth / nnormal;
          }
          share = share < room ? share : room;
          int got = 0;
          for (int prio = 0; prio < 2 && share > 0; prio++) {
            const int qi = 2 * p + prio;
            unsigned h = 0;
            int k = 0;
            if (lane == 0) {
              for (;;) {
                h = ldv(fq_head(fq, qi));
                const int avail = (int)(ldv(fq_tail(fq, qi)) - h);
                k = avail < share ? avail : share;
                if (k <= 0) { k = 0; break; }
                if (atomicCAS(fq_head(fq, qi), h, h + (unsigned)k) == h) break;
              }
            }
            h = __shfl_sync(FULLMASK, h, 0);
            k = __shfl_sync(FULLMASK, k, 0);
#ifdef LCR_FLOW_DEBUG
            if (lane == 0 && k > 0) atomicAdd(fq.ctl + 992, (unsigned)k);
#endif
            const int t0q = __shfl_sync(FULLMASK, LCR_VOL(lq_tail), 0);
            if (lane < k) {
              // ticket h + lane: take the item (its producer may be between its atomicAdd and its store), free the slot
              unsigned long long* slot = fq_slot(fq, qi, h + (unsigned)lane);
              unsigned long long v = 0;
              long long t1 = 0;
              bool ok = true;
              while ((unsigned)((v = ldv64(slot)) >> 32) != h + (unsigned)lane + 1u) {
                const long long now = clock64();
                if (t1 == 0) t1 = now;
                else if (now - t1 > LCR_FLOW_WATCHDOG) { atomicExch(fq_error(fq), 2u); ok = false; break; }
              }
              if (ok) stv64(slot, (unsigned long long)(h + (unsigned)lane + fq.mask[qi] + 1u) << 32);
              FLOW_COUNT(fq, 33);
              LCR_VOL(lq_item[(t0q + lane) & (LCR_LQ - 1)]) = ok ? (unsigned)v : fq_item(0, 0, 0);
              LCR_VOL(lq_ph[(t0q + lane) & (LCR_LQ - 1)]) = ok ? p : -1;  // (-1: a hole; the kernel is ending with an error anyway)
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0 && k > 0) LCR_VOL(lq_tail) = t0q + k;  // publish the entries to the CTA
            __syncwarp();
            share -= k;
            got += k;
          }
          if (got > 0) {
            if (lane == 0) LCR_VOL(role) = p;
            break;
          }
          continue;  // somebody else was faster: look again
        }
        if (lane == 0) {
          stop = ldv(fq_remaining(fq)) == 0 || ldv(fq_error(fq)) != 0;
          const long long now = clock64();
          if (idle0 == 0) idle0 = now;
          else if (now - idle0 > LCR_FLOW_WATCHDOG) {
            if (atomicCAS(fq_error(fq), 0u, 1u) == 0u) {  // debug words: where the unfinished envs' items sit (tail - head per phase; BIG in the high half of [5])
              unsigned* dbg = fq_error(fq) + 1;
              for (int ph2 = 0; ph2 < FQ_NPH; ph2++) {
                const unsigned dd = (ldv(fq_tail(fq, 2 * ph2)) - ldv(fq_head(fq, 2 * ph2))) + (ldv(fq_tail(fq, 2 * ph2 + 1)) - ldv(fq_head(fq, 2 * ph2 + 1)));
                if (ph2 < 6) dbg[ph2] = dd; else dbg[5] |= dd << 16;
              }
            }
            stop = 1;
          }
        }
        stop = __shfl_sync(FULLMASK, stop, 0);
        if (stop) break;
        __nanosleep(backoff);
        backoff = backoff < 8192 ? backoff * 2 : 8192;
      }
      if (lane == 0) {
        if (stop) LCR_VOL(quit) = 1;
        atomicSub(&n_idle, 1);
        __threadfence_block();
        atomicExch(&scout_lock, 0);
      }
      __syncwarp();
      if (stop) break;
    }
#undef LCR_VOL
  }
  if (stats && lane == 0) {
    LCR_FLOW_TICK(7);
    for (int k = 0; k < 8; k++) if (tb[k]) atomicAdd(stats + k, (unsigned long long)tb[k]);
  }
#undef LCR_FLOW_TICK
}

}  // namespace lcr
