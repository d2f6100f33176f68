/*
 * mas_b200.h -- C ABI of the B200 batched iLQR engine (libmas_b200.so).
 *
 * Drop-in boundary for the iLQR path of markomiz/multi_agent_solver.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repository root).  Plain
 * pointers and sizes only; all host arrays are caller-owned, row-major as written below, which for a
 * single problem is exactly the memory of the reference's column-major Eigen matrices
 * (StateTrajectory n x (T+1) column-major == [T+1][n] row-major; ControlTrajectory m x T == [T][m]).
 *
 * The reference passes problems as std::function callbacks (types.hpp:21-50), which cannot run on
 * a GPU; here a problem is `model_id` + a POD parameter block selecting a registered device functor
 * set (multi_agent_solver_b200/csrc/models.cuh), and `deriv_mask` states which derivative callbacks
 * are analytic -- the rest use the finite-difference defaults OCP::initialize_problem() installs
 * (ocp.hpp:117-135).
 *
 * Every function returns MAS_B200_OK or an error code; mas_b200_last_error() gives the message of
 * the calling thread's last failure.  There is no CPU fallback: without a CUDA device every call
 * that touches a context fails with MAS_B200_ERR_CUDA.
 */
#ifndef MAS_B200_H
#define MAS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes; the C++ facade rethrows them as the reference's exception types ------------- */
#define MAS_B200_OK 0
#define MAS_B200_ERR_INVALID_ARGUMENT 1 /* std::invalid_argument (examples/example_utils.hpp:77-110) */
#define MAS_B200_ERR_OUT_OF_RANGE 2     /* std::out_of_range, missing required solver param (solvers/ilqr.hpp:42-44) */
#define MAS_B200_ERR_CUDA 3             /* std::runtime_error */
#define MAS_B200_ERR_NCCL 4             /* std::runtime_error */
#define MAS_B200_ERR_UNSUPPORTED 5      /* std::runtime_error: feature of the reference not on the device path */

/* ---- registered models (the reference's example OCPs) ---------------------------------------- */
#define MAS_B200_MODEL_SINGLE_TRACK_LANE 0 /* examples/single_track_ocp.cpp:14-116         n4 m2 */
#define MAS_B200_MODEL_SINGLE_TRACK_CIRC 1 /* examples/multi_agent_single_track.cpp:31-72  n4 m2 */
#define MAS_B200_MODEL_LQR4 2              /* examples/multi_agent_lqr.cpp:21-76           n4 m4 */
#define MAS_B200_MODEL_PENDULUM 3          /* examples/pendulum_swing_up.cpp:29-117        n2 m1 */
#define MAS_B200_MODEL_ROCKET 4            /* examples/rocket_max_altitude.cpp:31-137      n3 m1 */
/* not a reference example: model 0 plus one equality and one inequality path constraint (OCP::equality_constraints /
 * inequality_constraints, ocp.hpp:59-66), the vehicle for iLQR's augmented-Lagrangian terms (ilqr.hpp:121-170,236-260) */
#define MAS_B200_MODEL_SINGLE_TRACK_LANE_CONSTRAINED 5
#define MAS_B200_NUM_MODELS 6

/* derivative-mode bits: set = analytic callback installed, clear = finite-difference default */
#define MAS_B200_DERIV_A (1u << 0)   /* OCP::dynamics_state_jacobian   (ocp.hpp:71)  */
#define MAS_B200_DERIV_B (1u << 1)   /* OCP::dynamics_control_jacobian (ocp.hpp:72)  */
#define MAS_B200_DERIV_LX (1u << 2)  /* OCP::cost_state_gradient       (ocp.hpp:73)  */
#define MAS_B200_DERIV_LU (1u << 3)  /* OCP::cost_control_gradient     (ocp.hpp:74)  */
#define MAS_B200_DERIV_LXX (1u << 4) /* OCP::cost_state_hessian        (ocp.hpp:75)  */
#define MAS_B200_DERIV_LUU (1u << 5) /* OCP::cost_control_hessian      (ocp.hpp:76)  */
#define MAS_B200_DERIV_LUX (1u << 6) /* OCP::cost_cross_term           (ocp.hpp:77)  */
#define MAS_B200_DERIV_VX (1u << 7)  /* OCP::terminal_cost_gradient    (ocp.hpp:78)  */
#define MAS_B200_DERIV_VXX (1u << 8) /* OCP::terminal_cost_hessian     (ocp.hpp:79)  */
/* models with path constraints: analytic constraint Jacobians (clear = the central-difference defaults of ocp.hpp:137-171) */
#define MAS_B200_DERIV_EQ_JX (1u << 9)    /* OCP::equality_constraints_state_jacobian     (ocp.hpp:65) */
#define MAS_B200_DERIV_EQ_JU (1u << 10)   /* OCP::equality_constraints_control_jacobian   (ocp.hpp:66) */
#define MAS_B200_DERIV_INEQ_JX (1u << 11) /* OCP::inequality_constraints_state_jacobian   (ocp.hpp:67) */
#define MAS_B200_DERIV_INEQ_JU (1u << 12) /* OCP::inequality_constraints_control_jacobian (ocp.hpp:68) */

/* per-problem result flag; the reference's solve() returns void, the definition is SURVEY 8a */
#define MAS_B200_STATUS_CONVERGED 0  /* break at solvers/ilqr.hpp:269-271 */
#define MAS_B200_STATUS_MAX_ITER 1   /* loop at :82 exhausted */
#define MAS_B200_STATUS_TIME_LIMIT 2 /* break at :85-90 */

/* strategies (strategies/strategy.hpp:13, examples/example_utils.hpp:94-110) */
#define MAS_B200_STRATEGY_CENTRALIZED 0
#define MAS_B200_STRATEGY_SEQUENTIAL 1
#define MAS_B200_STRATEGY_LINESEARCH 2
#define MAS_B200_STRATEGY_TRUSTREGION 3

#define MAS_B200_MAX_CONTROL_DIM 8
#define MAS_B200_MAX_PARAMS 8

typedef struct mas_b200_context* mas_b200_context_t; /* one device + one stream; single owner, not thread-safe */
typedef struct mas_b200_batch* mas_b200_batch_t;     /* a batch of same-shaped OCPs resident in HBM */

/* struct OCP as far as iLQR reads it (ocp.hpp:30-81): dims, horizon, dt, input bounds, which
 * derivative callbacks are analytic, and the cost/dynamics constants of the selected model. */
typedef struct {
  int model_id;
  int state_dim;     /* must equal the model's; checked */
  int control_dim;
  int horizon_steps; /* OCP::horizon_steps */
  double dt;         /* OCP::dt */
  unsigned deriv_mask;
  int has_input_bounds; /* both input_lower_bounds and input_upper_bounds set (ilqr.hpp:213) */
  double input_lower[MAS_B200_MAX_CONTROL_DIM];
  double input_upper[MAS_B200_MAX_CONTROL_DIM];
  int num_params; /* 0 = the example's constants */
  double params[MAS_B200_MAX_PARAMS];
} mas_b200_ocp_desc;

/* iLQR::set_params keys (solvers/ilqr.hpp:40-55); defaults of the constructor (:26-37) */
typedef struct {
  int max_iterations;
  double tolerance;
  double max_ms; /* wall-clock budget for the whole batch, checked before each iteration; INFINITY disables */
  int debug;
  double penalty;
  double penalty_increase;
  double constraint_tolerance;
  double inequality_activation_tolerance;
} mas_b200_ilqr_params;

/* Raw HBM view of a batch for callers that keep data resident (layout: DESIGN.md "HBM layout"). */
typedef struct {
  int batch, ld, state_dim, control_dim, horizon_steps;
  double* x0;     /* [n][ld] */
  double* X;      /* [T+1][n][ld]  best_states   */
  double* U;      /* [T][m][ld]    best_controls */
  double* cost;   /* [ld]          best_cost     */
  int* iterations;
  int* status;
  double* params; /* [np][ld] or NULL */
} mas_b200_device_view;

typedef struct {
  long long iterations;   /* sum over problems */
  long long alpha_trials; /* candidates the sequential reference would evaluate */
  long long reg_retries;
  long long kernel_launches; /* since batch creation */
  int outer_iterations_run;  /* host loop trips of the last solve */
  int forward_lanes, forward_chains;
} mas_b200_batch_stats;

/* Per-kernel device time of the solves run while profiling was enabled: CUDA events recorded on the
 * context stream around every launch (the kernels run back to back on that one stream). */
typedef struct {
  double prologue_ms, backward_ms, forward_ms;
  long long prologue_launches, backward_launches, forward_launches;
  long long problem_iterations; /* sum over iterations of the problems still active = units one backward+forward pair processed */
  long long solves;
} mas_b200_profile;

const char* mas_b200_last_error(void);
int mas_b200_version(void);

/* iLQR::iLQR() defaults (solvers/ilqr.hpp:26-37) */
void mas_b200_ilqr_default_params(mas_b200_ilqr_params* p);

/* Model registry: examples::make_* equivalents.  default_params may be NULL. */
int mas_b200_model_info(int model_id, int* state_dim, int* control_dim, int* num_params, unsigned* available_mask, unsigned* example_mask,
                        double* default_params);
/* The example's OCP as the reference main builds it (dims, T, dt, bounds, derivative mode). */
int mas_b200_example_desc(int model_id, mas_b200_ocp_desc* out);
/* The example's initial_controls, [T][m] (zeros except pendulum sinusoid / rocket half thrust). */
int mas_b200_example_controls(int model_id, int horizon_steps, double* U);

/* device_id < 0: current device.  stream: a cudaStream_t to run on, or NULL to create one. */
int mas_b200_context_create(int device_id, void* stream, mas_b200_context_t* out);
int mas_b200_context_destroy(mas_b200_context_t ctx);
int mas_b200_context_synchronize(mas_b200_context_t ctx);

/* How the host waits for the device in the solve loop and in mas_b200_batch_wait_solution: 0 (default) spins, the lowest
 * latency when every pipeline has a core of its own; 1 sleeps on the event (cudaEventBlockingSync) -- for hosts that drive
 * more pipelines than they have cores (8 GPUs x 8 solves in flight on 32 vCPUs).  Applies to batches created afterwards. */
int mas_b200_context_set_blocking_sync(mas_b200_context_t ctx, int enable);

/* Pinned host memory for callers that want full-speed copies. */
int mas_b200_host_alloc(size_t bytes, void** out);
int mas_b200_host_free(void* p);

/* ---- batch of OCPs: struct OCP x batch ---------------------------------------------------------- */
int mas_b200_batch_create(mas_b200_context_t ctx, const mas_b200_ocp_desc* desc, int batch, mas_b200_batch_t* out);
int mas_b200_batch_destroy(mas_b200_batch_t b);
/* OCP::initial_state, [batch][n] */
int mas_b200_batch_set_initial_states(mas_b200_batch_t b, const double* x0);
/* per-problem model constants, [batch][num_params]; NULL = desc.params for all */
int mas_b200_batch_set_params(mas_b200_batch_t b, const double* params);
/* OCP::initial_controls -> best_controls, [batch][T][m]; NULL = zeros */
int mas_b200_batch_set_controls(mas_b200_batch_t b, const double* U);
/* OCP::initialize_problem (ocp.hpp:102-183): rollout of the controls and best_cost */
int mas_b200_batch_initialize(mas_b200_batch_t b);
/* mas::solve(Solver&, OCP&) (solvers/solver.hpp:28-32 -> iLQR::solve, solvers/ilqr.hpp:59-273) for
 * every problem of the batch; warm-starts from best_controls.  Asynchronous on the context stream. */
int mas_b200_batch_solve(mas_b200_batch_t b, const mas_b200_ilqr_params* params);
/* OCP::best_states [batch][T+1][n], best_controls [batch][T][m], best_cost [batch], plus the
 * counters the reference lacks.  Any pointer may be NULL.  Synchronises. */
int mas_b200_batch_get_solution(mas_b200_batch_t b, double* X, double* U, double* cost, int* iterations, int* status);
/* The same results without holding up the context stream for the PCIe transfer: the solution is staged in HBM on
 * the context stream, the device-to-host copies run on a second stream, and the call returns at once -- the next
 * mas_b200_batch_set_* / mas_b200_batch_solve may follow immediately.  The host buffers (pinned memory for a truly
 * asynchronous copy) are valid after mas_b200_batch_wait_solution; a second begin waits for the first. */
int mas_b200_batch_begin_get_solution(mas_b200_batch_t b, double* X, double* U, double* cost, int* iterations, int* status);
int mas_b200_batch_wait_solution(mas_b200_batch_t b);
/* Results streamed to the host WHILE the solve runs.  Registers page-locked host buffers (cudaHostAlloc /
 * cudaHostRegister; any may be NULL, shapes as in mas_b200_batch_get_solution) as the destination of every following
 * mas_b200_batch_solve: the reference's solve leaves its result in the caller's OCP (best_states, best_controls,
 * best_cost; solvers/ilqr.hpp:71-73), and so does this -- whenever problems leave the active set (stop test :269-271,
 * iteration cap, time budget :85-90) a kernel on a side stream writes their rows straight into the buffers over PCIe,
 * concurrently with the remaining iterations of the other problems, so that by the end of the solve only the last
 * finishers are still travelling.  mas_b200_batch_wait_solution fences the buffers; the next solve / set_controls of the
 * batch waits on the device for the previous export.  Buffers that are not page-locked: MAS_B200_ERR_INVALID_ARGUMENT.
 * All NULL: unregister.  Synchronises the context stream (a setup call, not a per-solve one). */
int mas_b200_batch_set_result_sink(mas_b200_batch_t b, double* X, double* U, double* cost, int* iterations, int* status);
int mas_b200_batch_get_device_view(mas_b200_batch_t b, mas_b200_device_view* out);
/* Constrained models only: the multipliers and the penalty parameter persist from solve to solve like the members
 * of a reference solver object (ilqr.hpp:331-338,415); this makes the next solve start from a fresh solver
 * (multipliers 0, penalty = params->penalty).  A new batch starts fresh. */
int mas_b200_batch_reset_solver_state(mas_b200_batch_t b);
int mas_b200_batch_get_stats(mas_b200_batch_t b, mas_b200_batch_stats* out);
int mas_b200_batch_set_profiling(mas_b200_batch_t b, int enable); /* resets the accumulated profile */
int mas_b200_batch_get_profile(mas_b200_batch_t b, mas_b200_profile* out);
/* Tuning of the line-search kernel: lanes per problem (1,2,4,8,16; 0 = auto from batch size) and
 * step sizes rolled out together per lane (1 or 2; 0 = auto). */
int mas_b200_batch_set_tuning(mas_b200_batch_t b, int forward_lanes, int forward_chains);
/* The line search can keep the trajectories of its trial rollouts in HBM ((n + m) * T doubles per resident lane, about
 * 0.3 GB) so that the accepted step is a copy, not a second pass over T dependent steps.  enable = 1 (default): in
 * the launches over small active sets (4..16 lanes per problem), where it cuts 35 % off the line search; 2: also in
 * the warp-cooperative kernel of large active sets (measured slower on B200: the extra HBM writes cost more than the
 * second rollout); 0: off.  Dropped silently when the allocation fails.  Results are identical in every mode. */
int mas_b200_batch_set_trial_store(mas_b200_batch_t b, int enable);
/* How the backward pass (solvers/ilqr.hpp:92-193) is mapped: 0 = auto; 1 = one thread per problem, derivatives and
 * Riccati step fused; 2 = the finite-difference stencil points of a step dealt out to eight lanes per problem;
 * 3 = time-parallel: the derivatives of ALL time steps (ilqr.hpp:106-113 is independent across t) evaluated at once by
 * one kernel, then the Riccati recursion alone as a second one.  Auto takes 3 for active sets of at most max_problems
 * problems (default 8192; 0 = keep) and for finite-difference-heavy derivative modes whenever the derivative blocks
 * fit 512 MB.  Results are bit-identical in every mode. */
int mas_b200_batch_set_backward_mode(mas_b200_batch_t b, int mode, int max_problems);
/* The per-iteration trace of the last solve, when it ran with params->debug != 0 (the reference prints these lines to
 * std::cout, ilqr.hpp:79-80,262-267; a batch records them instead): records[r][6] for r = 0 .. *n_records - 1 of one
 * problem -- r = 0: {initial cost, initial merit, nan, nan, nan, nan}; r = it: {cost, merit, d_merit, eq_violation,
 * ineq_violation, accepted step-size index or -1} after iteration it.  Synchronises. */
int mas_b200_batch_get_debug_trace(mas_b200_batch_t b, int problem, int max_records, double* records, int* n_records);
/* How many independent solves the caller keeps in flight on this device (other batches on other streams; default 1).
 * The automatic lane mappings trade work for latency: a small active set evaluates all ten step sizes of the line
 * search at once on up to 16 lanes per problem when the device would otherwise idle.  With n solves in flight the
 * device is not idle, and each batch sizes its mappings for 1/n of it.  Results do not depend on the hint. */
int mas_b200_batch_set_concurrency_hint(mas_b200_batch_t b, int solves_in_flight);
/* How the line search (solvers/ilqr.hpp:195-228) is scheduled: 0 = auto (by active-set size),
 * 1 = all step sizes concurrently on `forward_lanes` lanes per problem, 2 = compacted rounds of two
 * step sizes over the problems still searching, 3 = warp-cooperative (a warp owns 32 problems and deals
 * its lanes out to the (problem, step size) tasks still needed).  Results are bit-identical in every mode. */
int mas_b200_batch_set_line_search_mode(mas_b200_batch_t b, int mode);

/* One-shot: set_initial_states + set_controls + initialize + solve + get_solution on host buffers.
 * U is in/out (initial_controls in, best_controls out; NULL = zero initial controls, not returned).
 * The device batch behind this call stays with the context and is reused by the next call with the same description
 * and batch size (each call still starts from a fresh solver state); mas_b200_strategy_run shares it.  A call of
 * another shape replaces it, mas_b200_context_destroy frees it.
 */
int mas_b200_ilqr_solve_batch(mas_b200_context_t ctx, const mas_b200_ocp_desc* desc, const mas_b200_ilqr_params* params, int batch,
                              const double* x0, const double* model_params, double* U, double* X, double* cost, int* iterations, int* status);

/* mas_b200_batch_get_debug_trace for the batch behind the last mas_b200_ilqr_solve_batch of this context. */
int mas_b200_ilqr_last_debug_trace(mas_b200_context_t ctx, int problem, int max_records, double* records, int* n_records);

/* ---- multi-agent strategies: mas::solve(Strategy&, MultiAgentProblem&) (strategies/strategy.hpp:15-19)
 * on n_scenarios independent scenarios of n_agents agents each; agents have ids 0..n_agents-1 in
 * array order (already the id-sorted block order of MultiAgentProblem::compute_offsets,
 * multi_agent_problem.hpp:37-50).  Arrays are [scenario][agent][...].  U_init: every agent's
 * initial_controls / best_controls before the first round, NULL = zeros.  trace_* may be NULL:
 * [scenario][outer][agent] inner iteration counts / accepted flags / best_cost after the round.
 * MAS_B200_STRATEGY_CENTRALIZED: the stacked problem always starts from zero controls (build_global_ocp never sets
 * initial_controls, multi_agent_problem.hpp:52-127) -- U_init is ignored; max_outer has no meaning, and the
 * iteration count of the stacked solve is written to trace_iterations[scenario][0][0] only when max_outer >= 1.
 * Stacks above 256 states (n_agents * state_dim) run the general stacked solve that mas_b200_strategy_run_mixed uses for
 * agents of different models: same results, workspace in HBM instead of shared memory. */
int mas_b200_strategy_run(mas_b200_context_t ctx, int strategy, const mas_b200_ocp_desc* agent_desc, const mas_b200_ilqr_params* params,
                          int max_outer, int n_scenarios, int n_agents, const double* x0, const double* model_params, const double* U_init,
                          double* X, double* U, double* costs, double* total_cost, int* trace_iterations, int* trace_accepted, double* trace_cost);

/* The same for agents of DIFFERENT models and shapes (MultiAgentProblem accepts any mix: compute_offsets /
 * build_global_ocp, multi_agent_problem.hpp:37-127; the reference's own test stacks a 2x1 and a 1x2 agent,
 * tests/ocp_tests.cpp:76-154): one description per agent; the per-agent arrays are given as arrays of n_agents
 * pointers, agent a's array holding n_scenarios entries of ITS shape (x0[a]: [scenario][n_a], X[a]: [scenario][T_a+1][n_a],
 * U[a]: [scenario][T_a][m_a], costs[a]: [scenario], model_params[a]: [scenario][np_a] or NULL, U_init[a] or NULL).
 * Nash strategies (sequential, line search, trust region): agents of one description share a device batch, the
 * groups advance round by round in lockstep, the line-search strategy's joint cost is summed over all agents in
 * index order.  trace_iterations (optional): [scenario][outer][agent].
 * Centralized over a mix (strategies/centralized.hpp:18-38 on build_global_ocp of the mix): one stacked iLQR solve per
 * scenario with run-time block shapes, every derivative by finite differences, horizon and dt of the FIRST agent, bounds
 * only when every agent has both, zero initial controls (U_init is not read).  The results keep the stacked horizon T_0:
 * X[a] is [scenario][T_0+1][n_a], U[a] is [scenario][T_0][m_a], costs[a] the agent's own objective on its rows, total_cost
 * the stacked best_cost; the iteration count goes to trace_iterations[scenario][0][0] when max_outer >= 1. */
int mas_b200_strategy_run_mixed(mas_b200_context_t ctx, int strategy, const mas_b200_ocp_desc* agent_descs, const mas_b200_ilqr_params* params,
                                int max_outer, int n_scenarios, int n_agents, const double* const* x0, const double* const* model_params,
                                const double* const* U_init, double* const* X, double* const* U, double* const* costs, double* total_cost,
                                int* trace_iterations);

/* MultiAgentProblem::compute_offsets + build_global_ocp (multi_agent_problem.hpp:37-127) for agents of any mix of registered
 * models, evaluated on the device: blocks sorted by agent id (agent_ids NULL = input order); dims_out = {total state dim,
 * total control dim, horizon of the FIRST block, 1 if ALL agents have both input bounds}; dt of the first block; bounds_out
 * [2][total_u] (lower, upper) written only in that case; block_agent / state_offsets / control_offsets [n_agents] in block
 * order.  With X [total_x], U [total_u]: the stacked dynamics (block diagonal), stage cost at time_index and terminal cost,
 * each cost the sum of the agents' terms in block order starting from 0.0.  X = U = NULL: structure only (no device).
 * This is what tests/ocp_tests.cpp:76-154 of the reference checks for a 2x1 and a 1x2 agent. */
int mas_b200_global_ocp_eval_mixed(mas_b200_context_t ctx, const mas_b200_ocp_desc* agent_descs, const unsigned long long* agent_ids, int n_agents,
                                   const double* X, const double* U, int time_index, double* dynamics_out, double* stage_cost_out,
                                   double* terminal_cost_out, int* dims_out, double* dt_out, double* bounds_out, int* block_agent, int* state_offsets,
                                   int* control_offsets);

/* ---- multi-GPU (one process per GPU).  unique_id: 128 bytes from mas_b200_nccl_unique_id on rank 0,
 * distributed by the caller (torch.distributed / MPI / file). ------------------------------------- */
int mas_b200_nccl_unique_id(void* id128);
int mas_b200_context_init_nccl(mas_b200_context_t ctx, const void* id128, int rank, int world_size);
/* On a context with a communicator, the Nash strategies of mas_b200_strategy_run end every outer round
 * (strategies/nash.hpp:86-87,105-177,196-245) with an NCCL all-gather of every agent's (best_states, best_controls,
 * best_cost), so that every rank holds the joint trajectory set of all ranks.  Every rank must pass the same
 * n_scenarios, n_agents, max_outer, strategy and horizon (checked up front with a small all-gather; a mismatch is
 * MAS_B200_ERR_INVALID_ARGUMENT on every rank, never a hang) -- pad the last shard when the units do not divide.
 *   agents_sharded = 0 (default): the ranks hold different SCENARIOS; total_cost[s] is the local scenario's sum.
 *   agents_sharded = 1: the ranks hold different AGENTS of the same n_scenarios scenarios (rank-major = id order,
 *     e.g. 1,024 agents as 8 x 128); total_cost[s] is the sum over ALL ranks' agents in block order
 *     (collect_solution, nash.hpp:23-37), formed from the gathered costs, identical on every rank. */
int mas_b200_context_set_agent_sharding(mas_b200_context_t ctx, int agents_sharded);
/* The joint set of the last Nash strategy run (after its last round): [world][n_scenarios][n_agents][...] with the
 * per-agent shapes of mas_b200_strategy_run, rank-major.  Any pointer may be NULL.  Synchronises. */
int mas_b200_strategy_get_joint(mas_b200_context_t ctx, double* X_all, double* U_all, double* costs_all);
/* Device time spent in the per-round collectives of the last run (CUDA events around every exchange, summed),
 * the number of rounds, and the bytes every rank received per round. */
int mas_b200_strategy_get_exchange_stats(mas_b200_context_t ctx, double* collective_ms, int* rounds, long long* bytes_per_round);

/* Synthetic inputs of the headline batch: x0_i = (0, Y, psi, v) with Y~U(-2,2),
 * psi~U(-0.5,0.5), v~U(0,2) from std::mt19937_64(seed), drawn in that order, problem-major.
 * Host-only helper (no device needed); x0 is [batch][4]. */
int mas_b200_synthetic_single_track_x0(unsigned long long seed, int batch, double* x0);

/* fp64 pipe peak probe (DFMA throughput in TFLOP/s on the context device) for roofline reporting */
int mas_b200_probe_fp64_peak(mas_b200_context_t ctx, double* tflops);

/* Device self-test of the straight-line divisions the line search's rollout step uses (portable_math.h: div_spec,
 * div_const_spec) against the division instruction, on `pairs` generated operand pairs: counts[0] pairs checked,
 * counts[1] pairs div_spec accepted as exact, counts[2] of those that differ from a / b (must be 0), counts[3] /
 * counts[4] the same for div_const_spec with the divisors 2.5 and 6. */
int mas_b200_selftest_division(mas_b200_context_t ctx, unsigned long long seed, long long pairs, long long* counts);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H */
