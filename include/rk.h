/*
 * rk.h -- C ABI of the B200 rollout kernels ("rk") behind JaxBackpropPlanner.
 *
 * The reference (pyRDDLGym-jax v3.2) has no FFI: its hot path is reached through Python
 * closures that JAX traces and XLA compiles.  Each entry point below names the reference
 * function whose device work it replaces (paths relative to pyRDDLGym_jax/core/).
 *
 * Conventions: every function returns 0 on success, a negative RK_E_* code for an invalid
 * argument / program, or a positive cudaError_t.  The caller owns every device buffer; the
 * library owns only the opaque program handle.  All launches go to the caller's stream and
 * never synchronise.  Nothing here falls back to the CPU.
 *
 * Batched buffers are "fluent-major": a quantity with W words per rollout is stored as the
 * concatenation over fluents f of a [B][n_f] array (the dict of [B, *objs] tensors of
 * JaxRDDLSimState.fls, compiler.py:54-62, flattened), and [T] such blocks when it has a
 * step axis.  Numerical error bits follow compiler.py:861-888.
 */
#ifndef RK_H
#define RK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RK_E_BADARG   (-1)
#define RK_E_BADPROG  (-2)
#define RK_E_SMEM     (-3)
#define RK_E_NOMEM    (-4)

typedef struct rk_program rk_program;

typedef struct rk_sizes {
  int32_t kind;          /* 0 forward, 1 backward                                         */
  int32_t relaxed;       /* 1 = JaxRDDLCompilerWithGrad model, 0 = exact JaxRDDLCompiler  */
  int32_t S, A, N;       /* state / action-parameter / noise words per rollout(-step)     */
  int32_t n_mparams;     /* model parameter slots (logic.py:266 aux['params'])            */
  int32_t rpw, lanes;    /* rollouts per warp, lanes per rollout                          */
  int32_t warp_words, const_words;
  int32_t n_prologue, n_step, n_epilogue;
  int32_t warps_per_cta; /* launch geometry chosen for this program                       */
  int32_t smem_bytes;
  int32_t static_err;
  int32_t tape_words;    /* words per rollout-step of the tape: S, or S + taped forward values */
} rk_sizes;

/* Upload a program produced by core/program.py to the current device.
 * Replaces: jax.jit tracing/compilation of compile_rollouts (compiler.py:621-666,
 * planner.py:2687, 2930). */
int rk_program_create(const void* blob, size_t nbytes, rk_program** out);
void rk_program_destroy(rk_program* prog);
int rk_program_query(const rk_program* prog, rk_sizes* out);
/* Attach a per-program specialised kernel (a cubin compiled from the CUDA C that
 * core/codegen.py generates out of the same instruction stream).  Subsequent launches of the
 * program use it instead of the generic interpreter kernel; semantics are identical.
 * Plays the role of the XLA executable cache behind jax.jit (planner.py:2687, 2930). */
int rk_program_attach_cubin(rk_program* prog, const void* image, size_t nbytes,
                            const char* entry);
/* The same with the launch geometry of the kernel given by the caller: threads per CTA (a multiple
 * of 32; 0 = the interpreter layout's warps_per_cta * 32), dynamic shared memory in bytes (-1 = the
 * interpreter layout's) and the preferred shared-memory carve-out in percent (-1 = driver default).
 * Thread-per-rollout kernels (core/codegen_tpr.py: lanes = 1, a warp carries 32 rollouts) keep all
 * their values in registers and attach with (128, 0, 0). */
int rk_program_attach_cubin_ex(rk_program* prog, const void* image, size_t nbytes,
                               const char* entry, int threads_per_cta, int smem_bytes, int carveout);
/* number of warps (= gradient partial rows) a launch with batch B uses */
int rk_program_num_warps(const rk_program* prog, int B);

/* Forward rollout of B trajectories for T steps.
 * Replaces: the jitted train_rollouts / test_rollouts executable, i.e. lax.scan over
 * _jax_wrapped_batched_policy_step (compiler.py:561-619).
 *   s0       [S*B]      initial state, fluent-major (init_sim_state, compiler.py:729-771)
 *   policy   [T*A]      straight-line plan parameters (planner.py:786-868), or NULL
 *   actions  [T*A*B]    per-rollout actions for externally computed policies, or NULL
 *   noise    [T*N*B]    supplied draws, or NULL when N == 0
 *   mparams  [n_mparams] relaxation weights, or NULL for the compiled defaults
 *   disc     [T]        gamma^t, or NULL when gamma == 1 (planner.py:2754-2757)
 *   reward   [T*B]      log['reward'] transposed (compiler.py:538)
 *   returns  [B]        sum_t gamma^t r_t (planner.py:2761)
 *   err      [T*B]      log['error'] bitmask (compiler.py:539), or NULL
 *   tape     [T*tape_words*B] state at the start of every step (+ the forward values the adjoint
 *                       program reads back when it was built with a full tape), or NULL
 *   final    [S*B]      state after the last step, or NULL
 *   traj     [T*TRAJ*B] logged fluents (cache_path_info, compiler.py:515-518), or NULL */
int rk_rollout_forward(const rk_program* prog, void* stream, int B, int T,
                       const void* s0, const float* policy, const void* actions,
                       const float* noise, const float* mparams, const float* disc,
                       float* reward, float* returns, uint32_t* err,
                       void* tape, void* final_state, void* traj);

/* Backprop-through-time of  sum_b gret[b] * return_b  w.r.t. the policy parameters.
 * Replaces: the reverse pass of jax.value_and_grad over the scan (planner.py:2873-2882).
 *   tape, noise, policy/actions, mparams, disc : as given to / written by the forward pass
 *   gret     [B]        d loss / d return_b  (-1/B for utility='mean', planner.py:2817-2818)
 *   gradp    [num_warps*T*A] per-warp partial gradients of the plan parameters (SLP), or NULL
 *   gact     [T*A*B]    per-rollout action gradients (external policies), or NULL
 *   gs0      [S*B]      gradient w.r.t. the initial state, or NULL
 *   gs_in    [S*B]      cotangent of the state after the last step (programs built with
 *                       state_ct_in; NULL = zero).  With T = 1 launches this chains the adjoint
 *                       across decision epochs when a closed-loop policy (DRP) sits between
 *                       them. */
int rk_rollout_backward(const rk_program* prog, void* stream, int B, int T,
                        const void* tape, const float* policy, const void* actions,
                        const float* noise, const float* mparams, const float* disc,
                        const float* gret, float* gradp, float* gact, float* gs0,
                        const float* gs_in);

/* Fixed-order reduction of the per-warp gradient partials: grad[i] = sum_w gradp[w][i].
 * Replaces: the batch reduction XLA fuses into the transposed scan. */
int rk_reduce_partials(void* stream, const float* gradp, int num_rows, int n, float* grad);

/* Optimiser step + box projection on a flat parameter vector.
 * Replaces: optax chain update + apply_updates + plan.projection (planner.py:2885-2899,
 * 2349-2387; box projection planner.py:878-910).
 *   kind: 0 sgd, 1 rmsprop, 2 adam.  hyper[9] = {lr, decay|b1, eps, b2, clip_norm(<=0 off),
 *   momentum, step(1-based), 1 - decay|b1, 1 - b2}: the two complements are computed by the caller in
 *   double precision and rounded once, as optax does with its Python-float decay rates.  state: rmsprop nu [n]; adam mu [n] then nu [n].
 *   lo/hi [n] box (may be +-inf) or NULL.  zero_flag: set to 1 when all |g| <= 1e-8
 *   (planner.py:2861-2864). */
int rk_optimizer_step(void* stream, int kind, const float* hyper, int n,
                      float* params, float* state, const float* grad,
                      const float* lo, const float* hi, int32_t* zero_flag);

/* Gradient reduction FUSED with the optimiser step: grad[i] = sum_w gradp[w][i] (the order of
 * rk_reduce_partials) and, in the same launch, the update of parameter i from it (rk_optimizer_step without
 * global-norm clipping, which needs the whole gradient first).  The tail of the adjoint and the optimiser
 * are one kernel: planner.py:2873-2899 (value_and_grad's batch reduction, optax update, apply_updates, box).
 *   counters [2] device int32, zeroed once by the caller: the CTAs agree on the all-zero flag through them
 *   (integer atomics, order-independent) and leave them zero again. */
int rk_reduce_optimizer_step(void* stream, const float* gradp, int num_rows, int n, float* grad, int kind,
                             const float* hyper, float* params, float* state, const float* lo,
                             const float* hi, int32_t* zero_flag, int32_t* counters);

/* The optimiser step with the optional links of the optax chain (planner.py:2349-2387, 2885-2899):
 * clip_by_global_norm -> add_noise -> optimiser -> reduce_on_plateau -> ema -> apply_updates -> box.
 * Replaces: optax.chain(...).update + optax.apply_updates for planners built with noise_kwargs,
 * reduce_on_plateau_kwargs or ema_decay.  kind / hyper / state / lo / hi / zero_flag as rk_optimizer_step.
 *   cfg [13] HOST floats: eta, gamma (add_noise: N(0, eta / t^gamma), t = 1-based step), ema_decay,
 *        factor, patience, rtol, atol, cooldown, accumulation_size, min_scale (reduce_on_plateau),
 *        use_noise, use_plateau, use_ema (0 / 1)
 *   count   [1] device: updates applied so far (incremented here)
 *   plateau [6] device: scale, best_value, plateau_count, cooldown_count, count, avg_value (or NULL)
 *   ema     [n] device: moving average of the updates (or NULL)
 *   loss    [1] device: the value reduce_on_plateau tests (the train loss of this gradient; or NULL)
 *   seed        key of the counter-based noise (Philox4x32-10, counter = (element quad, t))
 *   scratch [8] device floats (caller-owned; only read between the two launches of long vectors) */
int rk_optimizer_chain_step(void* stream, int kind, const float* hyper, int n, float* params,
                            float* state, const float* grad, const float* lo, const float* hi,
                            int32_t* zero_flag, const float* cfg, float* count, float* plateau,
                            float* ema, const float* loss, unsigned long long seed, float* scratch);

/* Concurrency projection of the boolean plan parameters onto max-nondef-actions, run after
 * rk_optimizer_step's box projection.
 * Replaces: _jax_wrapped_slp_project_to_max_constraint (planner.py:947-954), i.e. the
 * per-time-step JaxSogbofaActionProjection (method 0, planner.py:530-617) or
 * JaxSortingActionProjection (method 1, planner.py:489-527) vmapped over the horizon.
 *   params [T*A] plan parameters (in place); the boolean action fluents occupy the column
 *   ranges seg_off[i] .. seg_off[i]+seg_len[i] (HOST arrays, n_seg <= RK_MAX_BOOL_FLUENTS,
 *   in action-fluent order); seg_noop[i] = 1 when the fluent's default is true;
 *   seg_weight[i] = the sigmoid weight hyper-parameter of that fluent (planner.py:743).
 *   allowed = rddl.max_allowed_actions; wrap_sigmoid / min_prob as in JaxStraightLinePlan.
 *   converged [1]: cleared to 0 when some step still has surplus after max_iter (SOGBOFA);
 *   the caller sets it to 1 beforehand. */
#define RK_MAX_BOOL_FLUENTS 16
int rk_project_slp(void* stream, int method, int T, int A, float* params, int n_seg,
                   const int32_t* seg_off, const int32_t* seg_len, const int32_t* seg_noop,
                   const float* seg_weight, int allowed, int max_iter, int wrap_sigmoid,
                   float min_prob, int32_t* converged);

/* ---- deep reactive policy (JaxDeepReactivePolicy, planner.py:1122-1534) ------------------
 * The flax MLP (planner.py:1342-1449) runs between two T = 1 launches of the rollout kernel.
 *
 * rk_sgemm: C[M][N] = epi(op(A)[M][K] * B[K][N]), fp32, row-major.
 *   flags: RK_GEMM_A_TRANS (A stored [K][M]), RK_GEMM_ACCUMULATE (C_in + A*B, before the
 *   epilogue: lets a layer whose input is several fluent blocks be summed block by block), and the epilogue
 *   in bits 4..7: RK_EPI_NONE, RK_EPI_BIAS (+ bias[n]), RK_EPI_BIAS_TANH (nn.Dense + tanh,
 *   planner.py:1359-1361), RK_EPI_DTANH (* (1 - aux[m][n]^2): the tanh adjoint).
 *   split_k > 1 runs a deterministic split-K through `workspace` ([split_k][M][N] floats);
 *   split_k < 0 lets the library pick up to -split_k slices (enough CTAs for two waves).
 * Replaces: nn.Dense forward (planner.py:1360, 1376, 1398, 1407) and its transposes in the
 * reverse pass. */
#define RK_GEMM_A_TRANS    1
#define RK_GEMM_ACCUMULATE 2
#define RK_EPI_NONE        (0 << 4)
#define RK_EPI_BIAS        (1 << 4)
#define RK_EPI_BIAS_TANH   (2 << 4)
#define RK_EPI_DTANH       (3 << 4)
int rk_sgemm(void* stream, int flags, int M, int N, int K, const float* A, int lda,
             const float* B, int ldb, float* C, int ldc, const float* bias, const float* aux,
             int ldaux, float* workspace, int split_k);

/* First dense layer read straight from the fluent-major state (seg_off / seg_len: the state
 * fluents' word ranges, HOST arrays as for the action heads): out[B][H] = tanh(x W0 + b0); when
 * xp != NULL also packs xp[B][D+1] = [x | 1], the operand of the fused dW0 / db0 product.
 * Replaces: DRPStateLayer + the first nn.Dense (planner.py:1287-1319, 1358-1361). */
int rk_drp_layer0(void* stream, int B, int D, int H, int n_seg, const int32_t* seg_off,
                  const int32_t* seg_len, const float* x, const float* W0, const float* b0,
                  float* out, float* xp, int act);
/* Bias gradients: y[N] (+)= sum_b G[b][n], deterministic two-stage sum through
 * workspace[chunks][N]. */
int rk_colsum(void* stream, int B, int N, const float* G, int ldg, float* y, int accumulate,
              float* workspace, int chunks);
/* Fused backward of the action heads (NH <= 16, hl <= 1024), Hl and G read once:
 *   gz[b][c] = (sum_j G[b][j] Wh[c][j]) (1 - Hl[b][c]^2);
 *   grad_tail = [db_last (hl) | dWh (hl x NH) | dbh (NH)] += batch sums (the three blocks are
 *   contiguous in the flat parameter vector).  workspace: rk_drp_head_backward_ctas(B) rows of
 *   hl + hl NH + NH floats; partials are added in CTA order (deterministic).
 * Replaces: the VJP of the head nn.Dense + the last tanh (planner.py:1397-1421). */
int rk_drp_head_backward_ctas(int B);
int rk_drp_head_backward(void* stream, int B, int hl, int NH, const float* Hl, const float* G,
                         const float* Wh, float* gz, float* grad_tail, float* workspace, int act);
/* Backward of the first layer (D <= 16) in two passes over gz0 [B][H]:
 *   grad_W0b0 = [dW0 (D x H) | db0 (H)] += batch sums of x^T gz0 and gz0;
 *   gs (fluent-major state cotangent) += gz0 W0T, W0T = the transposed kernel [H][D].
 * workspace: rk_drp_head_backward_ctas(B) rows of (D + 1) H floats. */
int rk_drp_layer0_backward(void* stream, int B, int D, int H, int n_seg, const int32_t* seg_off,
                           const int32_t* seg_len, const float* x, const float* gz0,
                           const float* W0T, float* gs, float* grad_W0b0, float* workspace);
/* Input LayerNorm of the DRP over the non-boolean state words (planner.py:1300-1319; flax
 * nn.LayerNorm: y = (x - mean) rsqrt(var + eps) scale + bias, var = E[x^2] - E[x]^2).  x, y, gy, gx
 * are fluent-major state buffers; seg_norm[f] = 0 passes fluent f through, g >= 1 normalises it
 * together with the other fluents of group g (groups numbered 1..G: `normalize` = one group of all
 * non-bool fluents, `normalize_per_layer` = one group per fluent); scale / bias [Dn] follow the
 * normalised words in fluent order.  backward: gx += LayerNorm^T gy and grad_scale_bias =
 * [dscale (Dn) | dbias (Dn)] += batch sums through workspace[rk_state_layernorm_ctas(B)][2 Dn]
 * (added in CTA order). */
int rk_state_layernorm_ctas(int B);
int rk_state_layernorm_forward(void* stream, int B, int D, int n_seg, const int32_t* seg_off,
                               const int32_t* seg_len, const int32_t* seg_norm, float eps,
                               const float* x, const float* scale, const float* bias, float* y);
int rk_state_layernorm_backward(void* stream, int B, int D, int n_seg, const int32_t* seg_off,
                                const int32_t* seg_len, const int32_t* seg_norm, float eps,
                                const float* x, const float* scale, const float* gy, float* gx,
                                float* grad_scale_bias, float* workspace);
/* Static min-max scaling of the policy input (StaticNormalizer, planner.py:284-338):
 * y = (x - lo[d]) / (hi[d] - lo[d]) per state word d (lo = 0, hi = 1 passes a word through);
 * backward: gx += gy / (hi[d] - lo[d]).  Fluent-major state buffers, lo / hi [D] on the device. */
int rk_state_scale_forward(void* stream, int B, int D, int n_seg, const int32_t* seg_off,
                           const int32_t* seg_len, const float* lo, const float* hi, const float* x,
                           float* y);
int rk_state_scale_backward(void* stream, int B, int D, int n_seg, const int32_t* seg_off,
                            const int32_t* seg_len, const float* lo, const float* hi,
                            const float* gy, float* gx);
/* Hidden-layer activation ids of the DRP kernels (planner.py:1128, 1358-1361): 0 tanh, 1 relu, 2 sigmoid,
 * 3 softplus, 4 elu, 5 leaky_relu(0.01) -- the jax.nn functions whose derivative is a function of the
 * output, so the adjoints need only the stored activation.  The id is a LAUNCH ARGUMENT of every kernel
 * with a fused activation: bits 8..15 of the `flags` word of rk_sgemm and of the `epi` word of rk_tcgemm
 * (RK_ACT(id)), the `act` argument of rk_drp_layer0 / rk_drp_head_backward / rk_film_backward /
 * rk_drp_mlp2_* / rk_act_apply.  The library keeps no process state: calls are re-entrant per stream. */
#define RK_ACT(id) ((id) << 8)
/* gelu (6, jax's default tanh approximation) and silu (7) are not monotone: their derivative needs the
 * pre-activation.  A layer with one of them is  z = X W + b  (bias-only epilogue, or rk_drp_layer0 with the
 * identity id 8)  followed by  y = act(z)  (rk_act_apply), both kept; an activation-adjoint epilogue with
 * id 6 / 7 reads z as its aux operand.  The fused per-epoch kernels (rk_drp_mlp2_*) take ids 0-5 only. */
int rk_act_apply(void* stream, int act, size_t n, const float* z, float* y);
/* FiLM time conditioning of a hidden layer (planner.py:1083-1092, 1362-1363):
 * out[b][c] = gamma[c] z[b][c] + beta[c], gamma / beta [H] = the two Dense(time embedding) vectors of
 * the decision epoch.  backward: g = cotangent of out, z = the tanh output it was applied to;
 * gz = g gamma (1 - z^2) (may alias g), grad_gamma_beta = [dgamma (H) | dbeta (H)] += batch sums
 * through workspace[rk_film_chunks(B)][2 H] (added in chunk order). */
int rk_film_chunks(int B);
int rk_film_forward(void* stream, int B, int H, const float* z, const float* gamma,
                    const float* beta, float* out);
int rk_film_backward(void* stream, int B, int H, const float* g, const float* z,
                     const float* gamma, float* gz, float* grad_gamma_beta, float* workspace, int act);
/* Skinny products with N <= 16 (action heads, state-cotangent blocks), A read exactly once:
 * out[b][n] = (accumulate ? out : 0) + sum_k A[b][k] W[k][n] (+ bias[n]). */
int rk_rowdot(void* stream, int accumulate, int B, int N, int K, const float* A, int lda,
              const float* W, int ldw, float* out, int ldo, const float* bias);

/* The hidden dense layers on the tensor cores (tcgen05 + TMEM + TMA), 3xTF32 split for fp32-grade
 * accuracy:  C[M][N] = epi(A[M][K] * Bt[N][K]^T).  A is plain fp32 (split into TF32-exact hi +
 * fp32 remainder lo inside the kernel); the weight operand is passed pre-split, Bt = Bt_hi +
 * Bt_lo as produced by rk_tf32_split (once per optimiser step).
 * Constraints: 32 <= N <= 256, N % 32 == 0, K % 32 == 0, leading dimensions % 4 == 0,
 * 16-byte aligned pointers.  epi: 0 none, 1 + bias, 2 tanh(+ bias), 3 * (1 - aux^2).
 * Replaces: nn.Dense + tanh of the hidden stack (planner.py:1358-1361) and its transpose. */
int rk_tf32_split(void* stream, const float* x, float* hi, float* lo, size_t n);
int rk_tcgemm(void* stream, int epi, int M, int N, int K, const float* A, int lda,
              const float* Bt_hi, const float* Bt_lo, int ldb, float* C, int ldc,
              const float* bias, const float* aux, int ldaux);

/* loss[0] = -U(returns) and gret[b] = -scale * dU/dreturn_b for the utilities with a closed-form cotangent
 * (planner.py:2161-2247; the mean is rk_mean_loss): kind 1 mean_var, 2 mean_std, 3 mean_semivar,
 * 4 mean_semidev, 5 sharpe (p0 = risk_free, p1 = eps), 6 entropic / exponential, 7 the symlog-transformed
 * mean of use_symlog_reward (planner.py:2761-2763); p0 = beta otherwise.  One CTA, deterministic sums:
 * the planner iteration stays one CUDA graph.  loss or gret may be NULL. */
int rk_utility_loss(void* stream, int kind, int B, const float* returns, float p0, float p1, float scale,
                    float* loss, float* gret);

/* The same for a SHARDED batch: returns [B] is the all-gathered GLOBAL return vector (a utility
 * other than the mean is a function of all B returns, planner.py:2817-2818), gret [n_local] receives the
 * cotangents of this rank's rollouts [b0, b0 + n_local). */
int rk_utility_loss_shard(void* stream, int kind, int B, int b0, int n_local, const float* returns,
                          float p0, float p1, float scale, float* loss, float* gret);

/* The order-statistic utilities (planner.py:2195-2229) and their cotangents: kind 8 var (p0 = alpha),
 * 9 cvar (alpha), 10 cvar_ste (alpha), 11 expectile (alpha, max_iter fixed-point steps), 12 gini (p0 = gamma).
 * The return vector is sorted once on the device (radix sort of (return, rollout) pairs); percentiles
 * interpolate linearly like jnp.percentile.  returns [B] global vector, gret [n_local] for the rollouts
 * [b0, b0 + n_local); work: caller-owned device scratch of rk_utility_sorted_work_bytes(B) bytes.
 * Replaces: jit(utility) + its reverse-mode derivative inside value_and_grad (planner.py:2873). */
size_t rk_utility_sorted_work_bytes(int B);
int rk_utility_sorted_loss(void* stream, int kind, int B, int b0, int n_local, const float* returns,
                           float p0, int max_iter, float scale, float* loss, float* gret, void* work,
                           size_t work_bytes);

/* Returns and loss of a closed-loop (deep reactive policy) update from the per-step logs of its T
 * one-step rollout launches (planner.py:2747-2765, 2811-2819):
 *   returns[b] = sum_t disc[t] reward[t][b] (disc NULL = 1);  ent_scratch[b] = sum_t entropy[t][b];
 *   loss[0] = -mean(returns) - ent_coeff[0] * mean(ent_scratch)   (entropy NULL: no second term; loss NULL:
 *   returns only).  Deterministic (fixed-order) sums. */
int rk_drp_returns_loss(void* stream, int T, int B, const float* reward, const float* disc,
                        const float* entropy, const float* ent_coeff, float* returns, float* ent_scratch,
                        float* loss);

/* W [r][c] -> WT [c][r] (or NULL) and / or its TF32 hi / lo split WT_hi, WT_lo [c][r] (both or neither):
 * the operand forms of a Dense kernel the policy kernels read, rebuilt once per optimiser step. */
int rk_transpose_split(void* stream, int r, int c, const float* W, float* WT, float* WT_hi, float* WT_lo);

/* Supplied-noise tensor of a program from a counter-based generator (csrc/rk_optim.cu): one launch
 * fills out[T][N * B] (draw site s = block [B][n_s] at word offset offs[s] * B of every step's row;
 * offs[n_sites] = N) with Philox4x32-10 words (key = seed, counter = (quad of the row, step, draws[0])) through the
 * site's law -- kinds: 0 uniform, 1 normal, 2 exponential, 3 gumbel, 4 logistic, 5 laplace, 6 cauchy,
 * 7 standard Gamma (Marsaglia-Tsang rejection on the element's own Philox stream; conc [N] DEVICE floats =
 * the concentration of every noise word of a rollout, read at the gamma sites; NULL without such sites).
 * draws: device word holding the draw number (advance it with rk_noise_advance: a captured graph then
 * draws fresh noise at every replay), or NULL = 0.  kinds / offs are HOST arrays.
 * Replaces: the jax.random draws of a step's key (compiler.py:1627-2240) on the throughput path; the
 * parity tests keep supplying the noise tensor. */
int rk_draw_noise(void* stream, int T, int B, int n_sites, const int32_t* kinds, const int32_t* offs,
                  unsigned long long seed, const unsigned long long* draws, float* out, const float* conc);
int rk_noise_advance(void* stream, unsigned long long* draws);

/* Fused forward of a policy network with two hidden layers for one decision epoch:
 *   heads[B][NH] = act(act(x W0 + b0) W1 + b1) Wh + bh
 * Replaces: DRPStateLayer + the two Dense/activation layers + the per-action Dense heads of
 * JaxDeepReactivePolicy (planner.py:1287-1449) inside the scan body (compiler.py:552-558).
 *   act    : RK_ACT_* id of the hidden activation (an argument, not process state)
 *   x_fm   : the state in the rollout kernels' fluent-major layout (fluent f = block [B][len_f] at word
 *            offset off_f * B); D = sum of len_f <= 16 inputs; H0n, H1n multiples of 16, <= 256
 *   fwd_img: operand image of W0 / Wh built by rk_drp_mlp2_prepare (once per optimiser step),
 *            rk_drp_mlp2_image_floats(D, H0n, H1n, NH, 0) floats, 16-byte aligned
 *   b0 [H0n]; W1T_hi / W1T_lo [H1n][H0n]: W1^T split by rk_tf32_split; b1 [H1n]; bh [NH], NH <= 16
 *   H0 [B][H0n], H1 [B][H1n]: hidden activations for the adjoint, or NULL (the exact test sweep
 *            writes nothing but the head outputs); 32-byte aligned
 * One launch, one CTA per <= 256 rollouts: layer 0 is computed straight into the tensor-core operand
 * stages, the hidden layer runs on tcgen05 (3xTF32, both 128-row tiles share each weight tile), the
 * head product is a second tcgen05 product fed from the epilogue (csrc/rk_mlp2.cu). */
int rk_drp_mlp2_image_floats(int D, int H0n, int H1n, int NH, int backward);
/* profiling aid (tools/mlp2_phases.py): per-CTA clock64 stamps of the kernel phases, buf[grid][8]; NULL = off */
int rk_drp_mlp2_set_timing(unsigned long long* buf);
int rk_drp_mlp2_prepare(void* stream, int D, int H0n, int H1n, int NH, const float* W0,
                        const float* Wh, float* fwd_img, float* bwd_img);
int rk_drp_mlp2_forward(void* stream, int act, int B, int D, int H0n, int H1n, int NH, int n_seg,
                        const int32_t* seg_off, const int32_t* seg_len, const float* x_fm,
                        const float* fwd_img, const float* b0, const float* W1T_hi,
                        const float* W1T_lo, const float* b1, const float* bh, float* H0, float* H1,
                        float* heads);

/* Fused input-side adjoint of the same network for one decision epoch (the transpose of the scan
 * body's policy call in the reverse pass, planner.py:2873):
 *   gz1[B][H1n] = (gheads Wh^T) * act'(H1);  gz0[B][H0n] = (gz1 W1^T) * act'(H0);  gs += gz0 W0^T
 *   gheads [B][NH]: cotangent of the head outputs; H0 / H1: the activations rk_drp_mlp2_forward kept;
 *   bwd_img: operand image of W0 / Wh (rk_drp_mlp2_prepare; ..._image_floats(.., 1) floats);
 *   W1_hi / W1_lo [H0n][H1n]: W1 split by rk_tf32_split;
 *   gz0 / gz1: pre-activation cotangents, written once (operands of the weight-gradient products
 *   rk_drp_wgrad_*); gs: the fluent-major state cotangent, the input cotangent is ADDED to it.
 * Same pipeline and constraints as the forward. */
int rk_drp_mlp2_backward(void* stream, int act, int B, int D, int H0n, int H1n, int NH, int n_seg,
                         const int32_t* seg_off, const int32_t* seg_len, const float* gheads,
                         const float* H0, const float* H1, const float* bwd_img, const float* W1_hi,
                         const float* W1_lo, float* gz0, float* gz1, float* gs);

/* Weight gradients of the policy network as ONE reduction over all K = T * B (epoch, rollout) rows of
 * the tapes the two sweeps left in HBM (csrc/rk_wgrad.cu; the transposes of nn.Dense in the reverse
 * pass, planner.py:2873).  tcgen05 3xTF32, MN-major operands split in the kernel, one CTA per K-slice
 * holding the whole result in TMEM, partials added in CTA order (deterministic).
 *   hidden: dW [M][N] (+)= A[K][M]^T B[K][N], db [N] (+)= column sums of B (db may be NULL);
 *           32 <= M, N <= 256, multiples of 32; workspace rk_drp_wgrad_ctas() * (M * N + N) floats
 *   heads : dWh [M][NH] (+)= H[K][M]^T G[K][NH], dbh [NH] (+)= column sums of G (or NULL), NH <= 16;
 *           workspace ..._ctas() * (M + 1) * 16 floats
 *   input : dW0 [D][M] (+)= x^T gz0, db0 [M] (+)= column sums of gz0; gz0 [T * Bt][M]; x_tape: the
 *           fluent-major state tape, epoch t at word offset t * S_words * Bt, D <= 15 of its words in the
 *           seg_off / seg_len fluents; workspace ..._ctas() * M * 16 floats */
int rk_drp_wgrad_ctas(void);
int rk_drp_wgrad_hidden(void* stream, int accumulate, int M, int N, long long K, const float* A,
                        const float* B, float* dW, float* db, float* workspace);
int rk_drp_wgrad_heads(void* stream, int accumulate, int M, int NH, long long K, const float* H,
                       const float* G, float* dWh, float* dbh, float* workspace);
int rk_drp_wgrad_input(void* stream, int accumulate, int M, int D, int T, int Bt, int S_words,
                       int n_seg, const int32_t* seg_off, const int32_t* seg_len, const float* gz0,
                       const float* x_tape, float* dW0, float* db0, float* workspace);

/* Weight gradient of a hidden layer on the tensor cores: C[M][N] (+)= A[K][M]^T * B[K][N] with
 * both operands plain fp32 activations stored batch-major (K = rollouts), MN-major UMMA tiles,
 * 3xTF32 split done in the kernel, deterministic split-K (at most max_splits slices) through
 * workspace[max_splits][M][N].  Constraints: M % 32 == 0, 32 <= N <= 256, N % 32 == 0.
 * Replaces: the dW = X^T dZ transpose of nn.Dense in the reverse pass (planner.py:2873). */
int rk_tcgemm_tn(void* stream, int accumulate, int M, int N, int K, const float* A, int lda,
                 const float* B, int ldb, float* C, float* workspace, int max_splits);

/* Action heads of the DRP (planner.py:1366-1429) + clip to the action box (planner.py:1472-1475).
 *   heads [B][2A] (mu | log_sigma) when stochastic, [B][A] otherwise;
 *   seg_off / seg_len / seg_kind (HOST arrays, n_seg <= RK_MAX_ACTION_FLUENTS): the action
 *   fluents' word ranges, contiguous and in action-fluent order, and their kind: 0 real, 1 bool,
 *   2 int / enum (seg_kind NULL = all real); actions, noise and gact use the rollout kernels'
 *   fluent-major layout (fluent f = block [B][len_f] at word offset off_f * B);
 *   real / int: clip(mu + train * exp(clip(log_sigma)) * noise, lo, hi);
 *   bool: sigmoid(bool_weight * (logit + train * noise)), noise ~ Logistic (planner.py:1381-1391);
 *   train = 1 (train policy) or 0 (test policy, planner.py:1486); lo/hi [A] or NULL;
 *   typed bit 0: bool actions are written as int32 (p > 0.5) and int actions as int32 round(.),
 *   the input words of the exact test program (planner.py:1488-1497); typed bit 1: wrap_non_bool,
 *   real / int heads go through the sigmoid / softplus box wrap of planner.py:620-628 instead of
 *   the clip (the test policy applies both);
 *   entropy [B] = sum clip(log_sigma) - sum_bool [p log p + (1-p) log(1-p)], or NULL. */
#define RK_MAX_ACTION_FLUENTS 16
int rk_drp_actions_forward(void* stream, int B, int A, int stochastic, int n_seg,
                           const int32_t* seg_off, const int32_t* seg_len, const int32_t* seg_kind,
                           const float* heads, const float* noise, float train, const float* lo,
                           const float* hi, float ls_lo, float ls_hi, float bool_weight, int typed,
                           float* actions, float* entropy);
/* gheads [B][2A or A] = d actions / d heads applied to gact (clip: gradient 1 inside); boolean
 * logits also receive the gradient of the entropy regulariser, ent_coeff[0] * inv_batch * dH/dlogit
 * (ent_coeff: device scalar, NULL = none); wrap bit 0 differentiates the box wrap instead of the
 * clip, bit 1 = sigma_entropy_grad (planner.py:1414-1417: the log_sigma entropy keeps its gradient,
 * otherwise it is stop_gradient).  Fluents of kind 3 / 4 are left to rk_drp_topk_*. */
int rk_drp_actions_backward(void* stream, int B, int A, int stochastic, int n_seg,
                            const int32_t* seg_off, const int32_t* seg_len, const int32_t* seg_kind,
                            const float* heads, const float* noise, float train, const float* lo,
                            const float* hi, float ls_lo, float ls_hi, float bool_weight,
                            const float* ent_coeff, float inv_batch, int wrap, const float* gact,
                            float* gheads);

/* Pooled boolean heads under a concurrency constraint: the 'dense_topk' head + GumbelSoftmaxTopK of
 * planner.py:1027-1080, 1435-1447, 1456-1463.  Action fluents of seg_kind 3 (kind 4: default value
 * true, output flipped to 1 - a) share one head whose columns, in fluent order, are the pooled logits
 * (at most *max_pooled of them, allowed = max-nondef-actions <= *max_allowed: rk_drp_topk_limits).
 *   allowed == 1       : a = softmax(logits + train * gumbel)
 *   allowed > 1, train : allowed rounds of  l += safe_log(1 - khot);  khot += softmax(weight * l)
 *                        starting from l = logits + gumbel
 *   allowed > 1, test  : exact k-hot of the allowed largest logits
 * noise: Gumbel(0, 1) draws in the fluent-major action layout (NULL or stochastic == 0: none);
 * entropy[b] += -sum p log p with p = softmax(weight * logits) when stochastic; typed & 1 writes
 * (a > 0.5) as int32 words.  backward (train form): gheads columns of the pooled fluents =
 * d loss / d logits from gact and from the entropy term ent_coeff[0] * inv_batch (NULL = none); their
 * log_sigma columns are zeroed.  Launch after rk_drp_actions_forward / backward on the same buffers. */
int rk_drp_topk_limits(int* max_pooled, int* max_allowed);
int rk_drp_topk_forward(void* stream, int B, int A, int stochastic, int n_seg,
                        const int32_t* seg_off, const int32_t* seg_len, const int32_t* seg_kind,
                        const float* heads, const float* noise, int train, int allowed, float weight,
                        int typed, float* actions, float* entropy);
int rk_drp_topk_backward(void* stream, int B, int A, int stochastic, int n_seg,
                         const int32_t* seg_off, const int32_t* seg_len, const int32_t* seg_kind,
                         const float* heads, const float* noise, int allowed, float weight,
                         const float* ent_coeff, float inv_batch, const float* gact, float* gheads);

/* Tile the un-batched initial state (S words: float bits, or int32 0/1 / int words for the exact
 * model) into the fluent-major s0 buffer of B rollouts.  seg_off / seg_len: HOST arrays with the
 * state fluents' word ranges (n_seg <= RK_MAX_STATE_FLUENTS).
 * Replaces: the host-side np.tile of init_fls_and_nfls (compiler.py:668-727) and the transfer of
 * the tiled arrays: only S words cross PCIe per optimize() call / online-replanning step. */
#define RK_MAX_STATE_FLUENTS 64
int rk_tile_state(void* stream, int B, int n_seg, const int32_t* seg_off, const int32_t* seg_len,
                  const void* init, void* s0);

/* -mean(returns) and its cotangent, for utility='mean' (planner.py:2817-2818):
 * loss[0] = -(1/B) sum_b returns[b];  gret[b] = -1/B. */
int rk_mean_loss(void* stream, const float* returns, int B, float* loss, float* gret);

/* One-shot sum all-reduce of a short vector (n <= n_max floats) over NVLink peer memory, in
 * place, the same bits on every rank (sum in rank order).  peer_bufs[r] / peer_flags[r]: HOST
 * arrays of the device addresses, valid on THIS rank, of rank r's symmetric buffer (2 * world *
 * n_max floats) and flag words (2 * RK_MAX_PEERS uint32, zero at start-up); epoch_ctr: one device word
 * per rank, zero at start-up, advanced by the kernel (graph replay safe).  status (nullable) is
 * set to 1 if a peer does not answer within ~10 s.  Every rank must issue the same sequence of
 * calls.  Replaces: the batch mean inside value_and_grad (planner.py:2817-2818, 2873) once the
 * batch is sharded over GPUs. */
#define RK_MAX_PEERS 16
int rk_peer_allreduce(void* stream, int world, int rank, const uint64_t* peer_bufs,
                      const uint64_t* peer_flags, int n, int n_max, float* vec,
                      uint32_t* epoch_ctr, int32_t* status);

/* The same exchange with a gathered tail: the message is vec[0..n) followed by n_extra (<= 4) short
 * vectors read from their own buffers (extra_src[k], extra_len[k] floats); their sums are written to
 * extra_dst (contiguous).  One launch carries [gradient | train loss | test loss | loop-control votes]
 * of a planner iteration without staging copies (planner.py:3245-3262 reads exactly these). */
int rk_peer_allreduce_gather(void* stream, int world, int rank, const uint64_t* peer_bufs,
                             const uint64_t* peer_flags, int n, int n_max, float* vec,
                             uint32_t* epoch_ctr, int32_t* status, int n_extra,
                             const float* const* extra_src, const int32_t* extra_len, float* extra_dst);

#ifdef __cplusplus
}
#endif
#endif /* RK_H */
