/*
 * stochopy_b200 -- C ABI of the B200 population engine.
 *
 * The reference (keurfonluu/stochopy v2.3.0) is pure Python and has no FFI; the
 * entry points below are the functions a maintainer would bind (ctypes, see
 * INTEGRATION.md) to run the per-generation hot path of
 * stochopy.optimize.minimize() on a B200.  Each entry cites the reference code
 * it replaces.  All pointers named d_* / inside the state structs are DEVICE
 * pointers (row-major, leading dimension `ld` in elements, 16-byte aligned,
 * ld a multiple of 16/sizeof(T)); everything is enqueued on `stream`
 * (a cudaStream_t passed as void*, NULL = legacy default stream) and returns
 * immediately.  Return value: 0 on success, <0 on error (sp_last_error()).
 *
 * No function here computes on the CPU: if no CUDA device / kernel image is
 * usable the call fails.
 */
#ifndef STOCHOPY_B200_H
#define STOCHOPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_ABI_VERSION 2

/* error codes */
#define SP_OK 0
#define SP_ERR_ARG (-1)    /* bad argument (shape, enum, alignment) */
#define SP_ERR_CUDA (-2)   /* CUDA runtime error, text in sp_last_error() */
#define SP_ERR_SHAPE (-3)  /* ndim outside the compiled row shapes */

/* ctrl.status while the optimiser has not terminated */
#define SP_RUNNING (-1000)
/* ctrl.status when a kernel found its own state inconsistent (e.g. a ranking without rank 0): the
 * front-ends raise instead of returning a result */
#define SP_STATUS_INTERNAL (-902)

typedef enum { SP_F32 = 0, SP_F64 = 1 } sp_dtype;

/* stochopy/factory/benchmark.py; SP_OBJ_HOST = fitness supplied by the caller
 * (arbitrary Python fun(x): propose on device, evaluate on host, select on device) */
typedef enum {
  SP_OBJ_ACKLEY = 0,      /* benchmark.py:14-34  */
  SP_OBJ_GRIEWANK = 1,    /* benchmark.py:37-56  */
  SP_OBJ_QUARTIC = 2,     /* benchmark.py:59-76  */
  SP_OBJ_RASTRIGIN = 3,   /* benchmark.py:79-97  */
  SP_OBJ_ROSENBROCK = 4,  /* benchmark.py:100-118 */
  SP_OBJ_SPHERE = 5,      /* benchmark.py:121-136 */
  SP_OBJ_STYBLINSKI_TANG = 6, /* benchmark.py:139-156 */
  SP_OBJ_HOST = 100
} sp_objective;

/* de/_strategy.py:41-46 */
typedef enum { SP_DE_RAND1BIN = 0, SP_DE_RAND2BIN = 1, SP_DE_BEST1BIN = 2, SP_DE_BEST2BIN = 3 } sp_de_strategy;
/* de/_constraints.py:31-34, cpso/_constraints.py:69-72, cmaes/_constraints.py:85-87 */
typedef enum { SP_CONS_NONE = 0, SP_CONS_RANDOM = 1, SP_CONS_SHRINK = 2, SP_CONS_PENALIZE = 3 } sp_constraint;

/* Device-resident control block shared by every generation kernel of one
 * optimiser instance (64 bytes).  Kernels return at once when
 * status != SP_RUNNING, so generations can be enqueued ahead of the host. */
typedef struct {
  int32_t status;        /* SP_RUNNING or the reference's status code (_common.py:12-24) */
  int32_t nit;           /* generation that produced gbest/gfit */
  uint32_t done_blocks;  /* last-block-done counter, self-resetting */
  int32_t flag;          /* method specific (cpso: restart fired; es: hsig) */
  int64_t gbest_row;     /* row of the current best in the population */
  double gfit;           /* best objective value */
  double dist;           /* |gbest_old - gbest_new| of the last generation */
  double aux[3];         /* method specific (cpso: swarm radius, nw; es: sigma) */
} sp_ctrl;

/* ---- library ---------------------------------------------------------- */
int sp_abi_version(void);
const char* sp_last_error(void);
/* SM count, L2 bytes, total HBM bytes, compute capability (major*10+minor) */
int sp_device_info(int device, int* sm_count, int64_t* l2_bytes, int64_t* hbm_bytes, int* cc);
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t sp_launch_count(void);
/* scratch bytes every state struct needs behind `scratch` */
int64_t sp_scratch_bytes(void);

/* ---- a1+a2: batched objective  f[i] = fun(X[i] * scale + shift) ---------
 * replaces optimizer.wrapper (stochopy/optimize/_common.py:79-80) calling
 * stochopy/factory/benchmark.py once per row.  scale/shift may be NULL. */
int sp_eval(int objective, int dtype, const void* d_X, int64_t P, int N, int64_t ld,
            const void* d_scale, const void* d_shift, void* d_f, void* stream);

/* ---- a3: Latin hypercube init (stochopy/optimize/_common.py:109-120) -----
 * d_jitter (P x ld, U[0,1)) and d_perm (N x P int64, column permutations) may
 * both be NULL: then draws are Philox/Feistel from `seed` on the device. */
int sp_lhs_init(int dtype, void* d_X, int64_t P, int N, int64_t ld, const void* d_lower,
                const void* d_upper, uint64_t seed, const void* d_jitter, const int64_t* d_perm,
                void* stream);

/* LHS rows [row0, row0 + P_local) of a population of P_total (sharded swarm; Philox draws) */
int sp_lhs_init_shard(int dtype, void* d_X, int64_t P_local, int N, int64_t ld, const void* d_lower,
                      const void* d_upper, uint64_t seed, int64_t P_total, int64_t row0, void* stream);

/* ---- a4: synchronous selection (stochopy/optimize/_common.py:123-160) ----
 * rows with candfun < xfun: xfun = candfun and x = cand  (copy_when = 1), or,
 * for ping-pong populations, rows that do NOT improve are copied from `cand`
 * (the old population) into x (copy_when = 0).  Then argmin(xfun), distance to
 * the previous best, status ladder; result in ctrl / d_gbest. */
int sp_select_sync(int dtype, int it, int maxiter, double xtol, double ftol, const void* d_cand,
                   const void* d_candfun, void* d_x, void* d_xfun, int64_t P, int N, int64_t ld,
                   int copy_when, void* d_gbest, sp_ctrl* d_ctrl, void* d_scratch, void* stream);
/* first evaluation of a population: xfun given, gbest = x[argmin], no status */
int sp_best_init(int dtype, const void* d_x, const void* d_xfun, int64_t P, int N, int64_t ld,
                 void* d_gbest, sp_ctrl* d_ctrl, void* d_scratch, void* stream);

/* gbest exchange of a sharded swarm: d_recs = world records of rec_ld scalars,
 * [fit, x_0 .. x_{N-1}], all-gathered from the ranks; every rank picks the first minimum,
 * refreshes d_gbest and evaluates the status ladder exactly like selection_sync. */
int sp_gbest_reduce(int dtype, const void* d_recs, int world, int N, int64_t rec_ld, void* d_gbest, sp_ctrl* d_ctrl,
                    int it, int maxiter, double xtol, double ftol, void* stream);

/* ---- a6-a10: differential evolution (stochopy/optimize/de/_de.py:314-351,
 * de/_strategy.py, de/_constraints.py).  One call = one synchronous generation:
 * donors, mutation, binomial crossover, bound repair, objective, selection,
 * argmin and termination test, fused in one kernel. */
typedef struct {
  int32_t dtype, objective, strategy, constraint;
  int64_t P;
  int32_t N;
  int32_t maxiter;
  int64_t ld;
  double F, CR, xtol, ftol;
  uint64_t seed;
  void* X[2];      /* ping-pong populations; generation `it` reads X[it&1], writes X[(it&1)^1] */
  void* pbestfit;  /* (P) fitness of the population */
  void* pfit;      /* (P) candidate fitness of the last generation (_de.py:349: pfit = candfun) */
  void* gbest;     /* (ld) */
  const void* lower;
  const void* upper;
  sp_ctrl* ctrl;
  void* scratch;
  /* explicit draws (rng="numpy"); all NULL => Philox in-kernel */
  const void* r1;         /* (P x ld)  _de.py:250 */
  const int64_t* donors;  /* (k x P)   _de.py:306,311 first k rows */
  const int64_t* irand;   /* (P)       _de.py:340 */
  const void* repair;     /* (P x ld)  de/_constraints.py:24 uniform(lower,upper) */
} sp_de_state;
int sp_de_generation(const sp_de_state* st, int it, void* stream);
/* Chained generations.  SP_CHAIN_OUT: this generation leaves the reduction of its
 * per-CTA minima (gbest, |gbest_old - gbest_new|, nit, status: selection_sync,
 * _common.py:129-158) to the NEXT launch, which must carry SP_CHAIN_IN and resolves it
 * in its prologue before it touches the population (so nit/status are exactly the
 * reference's per-generation test).  ctrl/gbest are current only after a launch
 * without SP_CHAIN_OUT.  sp_de_chainable() != 0 when the state runs on the kernel
 * that supports this (in-kernel draws, device objective, full-warp rows). */
#define SP_CHAIN_IN 1
#define SP_CHAIN_OUT 2
int sp_de_chainable(const sp_de_state* st);
int sp_de_generation_chained(const sp_de_state* st, int it, int flags, void* stream);
/* propose only: writes the trial population U into X[(it&1)^1] (SP_OBJ_HOST path) */
int sp_de_propose(const sp_de_state* st, int it, void* stream);
/* enqueue generations it_first .. it_first+n-1 (Philox draws only) */
int sp_de_run(const sp_de_state* st, int it_first, int n, void* stream);

/* ---- a11-a14: PSO / competitive PSO (stochopy/optimize/cpso/_cpso.py:324-361,
 * 405-426, cpso/_constraints.py).  One call = velocity + position update with
 * Shrink, objective, personal-best selection, argmin, termination. */
typedef struct {
  int32_t dtype, objective, constraint, pad_;
  int64_t P;
  int32_t N;
  int32_t maxiter;
  int64_t ld;
  double w, c1, c2, xtol, ftol;
  double gamma, delta;  /* competitivity (<0: plain PSO) and swarm radius threshold */
  uint64_t seed;
  void* X;
  void* V;
  void* pbest;
  void* pbestfit;
  void* pfit;
  void* gbest;
  const void* lower;
  const void* upper;
  sp_ctrl* ctrl;
  void* scratch;
  const void* r1;  /* (P x ld) _cpso.py:262 */
  const void* r2;  /* (P x ld) _cpso.py:263 */
  /* one swarm sharded over GPUs (shard != 0): this rank holds rows [row0, row0 + P) of a
   * swarm of P_total; draws are keyed by the global row, so the run does not depend on the
   * GPU count.  The generation kernel then leaves gbest / status alone and writes its
   * local best [fit, x_0 .. x_{N-1}] into xch for the exchange (sp_gbest_reduce). */
  int64_t row0;
  int64_t P_total;
  void* xch;       /* (ld + 1) */
  int32_t shard;   /* 0 whole swarm; 1 sharded, exchange by the caller (xch + sp_gbest_reduce);
                    * 2 sharded, exchange inside the kernels over the peer mailboxes below */
  int32_t pad2_;
  /* shard == 2 (sp_peer_*): world mailboxes of sp_peer_bytes() each, one per rank, all mapped
   * into this process (CUDA IPC; NVLink peer memory on an NVSwitch box). */
  int32_t world, rank;
  void* mailbox;         /* this rank's own mailbox (device pointer) */
  void* const* peers;    /* DEVICE array [world]: rank r's mailbox as addressed from this process */
  /* optional, sp_pso_chain_scalars(ld) scalars: lets sp_pso_run chain the generations of a plain
   * PSO (gamma < 0, shard == 0) like sp_de_run does -- every CTA leaves its minimum and its best
   * row here and the next launch resolves gbest / status in its prologue */
  void* chain_rows;
} sp_pso_state;
int64_t sp_pso_chain_scalars(int64_t ld);
int sp_pso_generation(const sp_pso_state* st, int it, void* stream);
int sp_pso_propose(const sp_pso_state* st, int it, void* stream);
/* competitive restart (_cpso.py:405-426), split so host-drawn positions can be
 * injected: plan = swarm radius, nw = rows to reset (ctrl->flag, 0 if the swarm is
 * still wide), ascending stable rank of pbestfit into d_rank (P int32);
 * apply = V=0, X=pbest=uniform(lower,upper), pbestfit=1e30 for the nw worst.
 * d_fresh (nw x ld rows in the reference's reset order, worst first) may be NULL
 * (Philox keyed by particle row). */
int sp_cpso_restart_plan(const sp_pso_state* st, int it, int32_t* d_rank, void* stream);
int sp_cpso_restart_apply(const sp_pso_state* st, int it, const int32_t* d_rank, const void* d_fresh, void* stream);
int sp_cpso_restart(const sp_pso_state* st, int it, int32_t* d_rank, void* stream);
/* the same in pieces for a sharded swarm: local max |X_i - gbest|^2 into ctrl->aux[0]
 * (the caller max-reduces it over the ranks), then the decision from the global radius */
int sp_cpso_radius(const sp_pso_state* st, int it, void* stream);
int sp_cpso_decide(const sp_pso_state* st, int it, void* stream);
/* enqueue generations it_first .. it_first+n-1 (+ restart when gamma >= 0) */
int sp_pso_run(const sp_pso_state* st, int it_first, int n, int32_t* d_rank, void* stream);
/* CPSO with the restart taken out of the common path.  Whole swarm (shard == 0): per generation ONLY the
 * generation kernel is enqueued; its epilogue takes the reference's decision `radius < delta`
 * (_cpso.py:405-412) from a bound on the radius that needs no second pass over the swarm (the maximum
 * distance to the PREVIOUS gbest, accumulated in the row loop, +- the distance the gbest moved), which is
 * exact whenever the interval lies on one side of delta.  Peer-sharded swarm: generation kernel + a fused
 * radius + max-reduce + decision kernel.  When a restart fires (_cpso.py:405-426) the run parks with
 * ctrl.status = SP_STATUS_RESTART_PENDING (nit = the generation whose restart is due, ctrl.flag = nw, or
 * -1 if the bound could not decide) and the rest of the chunk returns at once; the host then calls
 * sp_cpso_restart_resume (exact radius decision if flag == -1, ranking, reset of the nw worst, status back
 * to SP_RUNNING) and continues at nit + 1. */
#define SP_STATUS_RESTART_PENDING (-901)
int sp_pso_run_lazy(const sp_pso_state* st, int it_first, int n, void* stream);
int sp_cpso_restart_resume(const sp_pso_state* st, int it, int32_t* d_rank, void* stream);

/* ---- 8e: one swarm row-sharded over GPUs, exchange fused into the kernels ----------
 * Reference analogue: the mpi backend's Bcast / Allreduce around the population
 * evaluation (_common.py:58-72); here the rows live on their GPUs and only the best
 * record moves.  Every rank owns a mailbox in device memory that all ranks map (CUDA IPC
 * handle exported by sp_peer_alloc, opened with sp_peer_open).  Per generation the last
 * CTA of the generation kernel STORES the rank's local best [fit, x_0..x_{N-1}] straight
 * into every peer's mailbox (NVLink stores), publishes a flag after a system-scope
 * fence, waits for the peers' flags and reduces the `world` records exactly like
 * selection_sync (first minimum, rank order = row order) -- gbest, dist, nit and status
 * come out identical on every rank with no NCCL call and no host round trip.  The
 * competitive restart exchanges the swarm radius (max) the same way inside the plan
 * kernel and, when it fires, the pbestfit shards (all-gather by peer stores) so every
 * rank ranks the whole swarm.  A peer that does not answer within ~10 s sets
 * ctrl.status = SP_STATUS_PEER_TIMEOUT on the waiting rank instead of hanging the GPU. */
#define SP_STATUS_PEER_TIMEOUT (-900)
#define SP_PEER_HANDLE_BYTES 64
int64_t sp_peer_bytes(int dtype, int world, int64_t ld, int64_t P_total);
/* cudaMalloc + zero-fill + export handle (SP_PEER_HANDLE_BYTES bytes, host memory) */
int sp_peer_alloc(int64_t bytes, void** d_mailbox, void* handle_out);
/* map a peer's mailbox from its handle; own mailbox: pass the pointer itself, no open needed */
int sp_peer_open(const void* handle, void** d_mailbox);
int sp_peer_close(void* d_mailbox);
int sp_peer_free(void* d_mailbox);
/* enqueue generations it_first .. it_first+n-1 of a shard == 2 swarm (+ restart when
 * gamma >= 0); d_rank_all: P_total int32 of scratch.  Every rank must enqueue the same calls. */
int sp_pso_run_sharded(const sp_pso_state* st, int it_first, int n, int32_t* d_rank_all, void* stream);


/* ---- 8f-3: user objectives compiled at run time (NVRTC) ------------------------------
 * The reference's fun(x, *args) -> float contract (_common.py:27-106) keeps arbitrary
 * Python callables on the host.  An objective given as CUDA C source that defines
 *     __device__ real objective(const real* x, int n)      (real = float or double)
 * is compiled for sm_100a and evaluated on the device: f[i] = objective(X[i] * scale + shift)
 * (scale/shift optional, _cmaes.py:168-173), one thread per individual on rows staged
 * through shared memory.  sp_jit_check only compiles (no GPU needed) and reports the
 * compiler log through sp_last_error(). */
int sp_jit_check(const char* source, int dtype, int64_t* cubin_bytes);
int sp_jit_compile(const char* source, int dtype, void** handle);
int sp_jit_eval(void* handle, int dtype, const void* d_X, int64_t P, int N, int64_t ld, const void* d_scale,
                const void* d_shift, void* d_f, void* stream);
int sp_jit_free(void* handle);

/* ---- counter-based draws as a buffer: out[row][j] = U[0,1) (normal = 0) or N(0,1)
 * (normal = 1) of Philox counter (j / (16/sizeof(T)), row, it, purpose); the same
 * streams the generation kernels consume in place (csrc/philox.cuh): LHS jitter (1), DE repair (4),
 * PSO restart (7), ES draws (8, 9, 10, 11), NA walks (12).  The DE crossover decisions (2) and the
 * PSO r1 / r2 coefficients (5, 6) are 16-bit pieces of differently keyed calls inside the kernels
 * (csrc/philox.cuh) and are not available here: those purposes return SP_ERR_ARG. */
int sp_random_fill(int dtype, void* d_out, int64_t P, int N, int64_t ld, int it, int purpose, uint64_t seed,
                   int normal, void* stream);

/* ---- fitness ranking (np.argsort(arfitness), _cmaes.py:272 / _vdcma.py:290) ----
 * d_rank[i] = number of individuals that sort before i (ascending, ties by
 * index = a stable argsort; NaN last); chunk sort + merge (csrc/rank.cuh), P < 2^31.
 * Scratch comes from the device's stream-ordered pool (cudaMallocAsync). */
int sp_fitness_rank(int dtype, const void* d_fit, int64_t P, int32_t* d_rank, void* stream);

/* ---- dense symmetric eigendecomposition (np.linalg.eigh, _cmaes.py:304) -------
 * d_C (N x N, ld = N) is symmetrised from its upper triangle in place
 * (_cmaes.py:303); eigenvalues ascending into d_w (N), eigenvectors into the
 * columns of d_B (N x N).  Cyclic one-sided Jacobi on the rows of W = Q0 C (Q0 = I or,
 * warm, the eigenvectors already in d_B); for the positive semi-definite covariance the
 * eigenpairs are lambda_j = |w_j|, q_j = w_j / |w_j| at convergence, so only W is rotated.
 * Every eigenvector is normalised so that its largest-magnitude component is positive
 * (LAPACK's sign is implementation defined).  N <= 64: one CTA, W in shared memory;
 * larger N: block Jacobi in one thread-block cluster of up to 16 CTAs (block pairs in
 * shared memory, hardware cluster barrier between outer rounds).
 * d_work: sp_sym_eigh_work_scalars(N) scalars. */
int64_t sp_sym_eigh_work_scalars(int N);
/* warm != 0: d_B holds an approximate eigenbasis to start from (few sweeps when C moved
 * little).  d_sweeps (optional, device int32): number of Jacobi sweeps performed. */
int sp_sym_eigh(int dtype, void* d_C, int N, void* d_w, void* d_B, void* d_work, int warm, int32_t* d_sweeps,
                void* stream);

/* ---- a15-a17: (mu,lambda)-CMA-ES generation (stochopy/optimize/cmaes/_cmaes.py:228-343,
 * converge :360-434, Penalize cmaes/_constraints.py:4-82).  Works in the space
 * standardised to [-1,1]; the objective sees x * xscale + xshift. */
typedef struct {
  sp_ctrl base;          /* status / nit / gfit (best of the last generation) / gbest_row */
  double sigma;          /* step size */
  double sigma_gen;      /* sigma the current population was sampled with */
  double ps_norm;
  double vd_ps;          /* vdcma: scalar evolution path */
  int64_t nfev;
  int64_t eigeneval;
  int32_t hsig;          /* cond of _cmaes.py:283 / _vdcma.py:303 */
  int32_t do_eig;        /* eigendecomposition due this generation */
  int32_t inject;        /* vdcma flg_injection */
  int32_t validfitval;   /* Penalize state */
  int32_t iniphase;
  int32_t hist_len;      /* Penalize dfithist length */
  int32_t sweeps;        /* Jacobi sweeps of the last decomposition */
  int32_t pad_;
  double aux[16];
} sp_es_ctrl;

typedef struct {
  int32_t dtype, objective, constraint, N;
  int64_t P, ld;
  int32_t mu, maxiter, ilim, hist_cap;
  double cc, cs, c1, cmu, damps, chind, mueff, xtol, ftol, insigma;
  uint64_t seed;
  void* xmean;     /* (N) */
  void* xold;      /* (N) */
  void* pc;        /* (N) */
  void* ps;        /* (N) */
  void* C;         /* (N x N) */
  void* B;         /* (N x N) eigenvectors in columns */
  void* D;         /* (N) sqrt of eigenvalues */
  void* invsqrtC;  /* (N x N) */
  void* arx;       /* (P x ld) population (standardised, unclipped) */
  void* arfit;     /* (P) fitness (penalised when constraint = Penalize) */
  void* Z;         /* (P x ld) N(0,I) draws of the generation; after the sampling GEMM its first mu rows hold
                    * the mu best individuals in rank order (panel of the mean / rank-mu update) */
  void* weights;   /* (mu) */
  void* xscale;    /* (ld) 0.5 (upper - lower) */
  void* xshift;    /* (ld) 0.5 (upper + lower) */
  void* besthist;  /* (maxiter) arbestfitness, zero initialised */
  void* work;      /* workspace, sp_cma_work_scalars(N, P) scalars */
  int32_t* rank;   /* (P) */
  void* bnd_weights; /* (N)      Penalize */
  void* dfithist;    /* (hist_cap) Penalize */
  sp_es_ctrl* ctrl;
  void* scratch;
  int32_t host_z;    /* 1: Z was filled by the caller (rng="numpy"), 0: Philox in-kernel */
  int32_t host_eigh; /* 1: the caller decomposes C itself when ctrl->do_eig (compat mode) */
} sp_cma_state;
int64_t sp_cma_work_scalars(int N, int64_t P);
/* one generation; with host_eigh the call stops after the covariance update when a
 * decomposition is due and sp_cma_finish_generation must follow once B, D are uploaded */
int sp_cma_generation(const sp_cma_state* st, int it, void* stream);
/* the generation in two halves around an external evaluation (host objectives):
 * sample = draws + sampling GEMM into arx; the caller fills arfit with fun(clip?(arx)
 * * xscale + xshift); update = penalty, selection, paths, covariance, eigenbasis, status */
int sp_cma_sample(const sp_cma_state* st, int it, void* stream);
int sp_cma_update(const sp_cma_state* st, int it, void* stream);
int sp_cma_finish_generation(const sp_cma_state* st, int it, void* stream);
int sp_cma_run(const sp_cma_state* st, int it_first, int n, void* stream);

/* ---- a18: VD-CMA generation (stochopy/optimize/vdcma/_vdcma.py:235-409): covariance
 * D (I + v v^T) D, O(N) per individual.  sample = injection vector, row-local
 * sampling y = D (z + (sqrt(1+|v|^2)-1)(z.vn) vn), x = xmean + sigma y, objective;
 * update = [Penalize], rank, weighted sums over the mu best (mean shift, pc, rank-mu
 * p/q vectors), sigma from the rank gap of the injected pair, natural-gradient step
 * on (v, D), termination ladder. */
typedef struct {
  int32_t dtype, objective, constraint, N;
  int64_t P, ld;
  int32_t mu, maxiter, ilim, hist_cap;
  double cc, c1, cmu, mueff, wsum, xtol, ftol, insigma;
  uint64_t seed;
  void* xmean;    /* (ld) */
  void* xold;     /* (ld) */
  void* dx;       /* (N) last mean shift */
  void* pc;       /* (N) */
  void* dvec;     /* (ld) */
  void* vvec;     /* (ld) */
  void* vn;       /* (ld) v / |v| (sp_vd_refresh) */
  void* diagC;    /* (N) dvec^2 (1 + vvec^2) of the current population */
  void* dy;       /* (ld) injected step */
  void* ginj;     /* (ld) N(0,I) vector of the injection (filled by the caller with host_z) */
  void* arx;      /* (P x ld) */
  void* ary;      /* (P x ld) y; holds the N(0,I) draws on entry with host_z */
  void* yvn;      /* (P) (y_i / dvec) . vn */
  void* arfit;    /* (P) */
  void* weights;  /* (mu) */
  void* xscale;
  void* xshift;
  void* besthist; /* (maxiter) */
  void* work;     /* sp_vd_work_scalars(N, P) scalars */
  int32_t* rank;
  void* bnd_weights;
  void* dfithist;
  sp_es_ctrl* ctrl;
  void* scratch;
  int32_t host_z;
  int32_t lean;   /* 1: the caller never reads arx (device objective, in-kernel draws, no Penalize): the
                   * sampling kernel stores y and the fitness only; x_i = xold + sigma_gen * y_i rebuilds a row */
} sp_vd_state;
int64_t sp_vd_work_scalars(int N, int64_t P);
/* |v|^2, vn, diagC from (vvec, dvec): once after initialisation */
int sp_vd_refresh(const sp_vd_state* st, void* stream);
int sp_vd_sample(const sp_vd_state* st, int it, int evaluate, void* stream);
int sp_vd_update(const sp_vd_state* st, int it, void* stream);
int sp_vd_generation(const sp_vd_state* st, int it, void* stream);
int sp_vd_run(const sp_vd_state* st, int it_first, int n, void* stream);

/* ---- a19: Neighbourhood Algorithm resampling (stochopy/optimize/na/_na.py:265-305).
 * The archive of every model ever evaluated lives on the device TRANSPOSED
 * (d_archT[j * cap + m], unit-cube coordinates) so the per-axis Voronoi scan reads
 * contiguous memory.  One CTA per new individual: Gibbs walk inside the cell of one
 * of the nr best archived models, axis after axis, limits by block reductions. */
/* append P rows (row-major, ld) at archive positions [M, M+P) */
int sp_na_append(int dtype, void* d_archT, int64_t cap, int64_t M, const void* d_X, int64_t P, int N, int64_t ld,
                 void* stream);
/* d_cells[r] = archive index of the model with fitness rank r, r < nr (d_rank from sp_fitness_rank) */
int sp_na_cells(const int32_t* d_rank, int64_t M, int32_t nr, int32_t* d_cells, void* stream);
/* d_X (P x ld) = new population.  d_u (P x ld) U[0,1) draws or NULL (Philox, seed/it);
 * d_mask (N int32, 0 = zero-span axis: coordinate fixed to 0 and skipped);
 * d_work: P * M scalars. */
int sp_na_resample(int dtype, const void* d_archT, int64_t cap, int64_t M, const int32_t* d_cells, int32_t nr,
                   void* d_X, int64_t P, int N, int64_t ld, const int32_t* d_mask, const void* d_u, uint64_t seed,
                   int it, void* d_work, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STOCHOPY_B200_H */
