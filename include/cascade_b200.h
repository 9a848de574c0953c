/* cascade_b200.h -- C ABI of libcascade_b200.so: batched, device-resident replacements for the
 * hot path of Cascade (MGessinger/Cascade).  Plain pointers and sizes only.
 *
 * Cascade has no plugin/FFI layer; its boundary is the public C API of libcascade.so
 * (reference src/acb_ode.h:25-67, src/cascade.h:8-26).  Each entry point below names the
 * reference function whose loop it runs for a whole batch; the scalar drop-in wrappers with the
 * reference's own names and signatures live in include/cascade_compat.h and call these with
 * batch = 1.  INTEGRATION.md shows the binding a Cascade maintainer would add.
 *
 * DATA LAYOUT (see include/cb_format.h).  prec = 32*L bits, L in {4, 32, 64, 128}.
 * A real ball is a "slot": L little-endian 32-bit mantissa limbs + one cb_meta record.
 * A complex ball is two adjacent slots (re, im).  A ball ARRAY of E entries with S slots per
 * entry is
 *      mant : uint32_t[E][L][S]     (limb-major inside an entry)
 *      meta : cb_meta [E][S]
 * with S = 2*batch (slot = 2*instance + part) for per-instance data and S = 2 for data shared by
 * every instance of the batch.
 *
 * "_dev" functions take DEVICE pointers, enqueue kernels on `stream` (a cudaStream_t, may be
 * NULL) and return without synchronising.  "_host" functions take HOST pointers in the same
 * layout, copy in, run, copy out and synchronise: they are the end-to-end path.
 *
 * Every function returns 0 on success, a CB200_E* code (< 0) for rejected arguments, or a
 * positive cudaError_t.  There is NO CPU fallback: without a usable CUDA device every compute
 * entry point fails with the CUDA error.
 */
#ifndef CASCADE_B200_H
#define CASCADE_B200_H

#include <stddef.h>
#include <stdint.h>
#include "cb_format.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_EINVAL (-1)     /* bad argument (unsupported L, negative sizes, ...) */
#define CB200_EVALUATION (-2) /* valuation >= 0: acb_ode_solve_fuchs is undefined there
                                 (reference src/fuchs_solver.c:108-136 never terminates) */
#define CB200_ERANGE (-3)     /* an integer factor b(b-1)...(b-i+1) would overflow int64 */
#define CB200_ENOMEM (-4)     /* workspace too small */

/* elementwise operation codes */
#define CB200_OP_ADD 0     /* acb_add */
#define CB200_OP_SUB 1     /* acb_sub        reference src/fuchs_solver.c:119 */
#define CB200_OP_MUL 2     /* acb_mul        reference src/fuchs_solver.c:118 */
#define CB200_OP_DIV 3     /* acb_div        reference src/fuchs_solver.c:137 */
#define CB200_OP_INV 4     /* acb_inv */
#define CB200_OP_MUL_SI 16 /* acb_mul_fmpz / acb_mul_si   reference src/fuchs_solver.c:130,135 */
#define CB200_OP_ADD_SI 17 /* acb_add_si     reference src/frobenius_solver.c:58 */

const char *cb200_version(void);
/* number of CUDA devices visible; < 0: CUDA error (negated) */
int cb200_device_count(void);
/* kernels launched by this library in this process so far (all entry points) */
uint64_t cb200_kernel_launches(void);

/* development aid: CB200_RR=1 in the environment makes the shared-operator recurrence try the reduced-radix
 * fused dot product (cascade_b200/csrc/cb_rr.cuh; same balls, measured slower: DESIGN.md section 4) before the
 * carry-chain one; with CB200_RR_STATS=1 as well, out[0] / out[1] = those products accepted / rejected so far */
int cb200_debug_rr_stats(uint64_t *out);

/* ---- workspace sizes (bytes of device memory the _dev entry points need) ----
 * Always ask these functions: what a workspace holds is the library's business and changes between builds (temporaries
 * of the reference's loops, the multiplier table and scheduling words of the recurrence kernel, at 2048 bits and up
 * the global-memory tails of the hybrid scratch columns of the Horner / Frobenius kernels).  The contents need no
 * initialisation and carry nothing from one call to the next.
 * Development switches read from the environment at launch time (no effect on results, only on which kernel form
 * runs): CB200_PERSISTENT=0/1 (work-unit recurrence kernel), CB200_HYBRID=0/1 (hybrid scratch columns),
 * CB200_TPB_CHOICE=0/1/2, CB200_RANGE=<coefficients per work unit>, CB200_COOP=0/1 (cooperating lanes: four per
 * instance; 2 / 4 force that many lanes for the Frobenius recurrence), CB200_FROBPERS=0/1 (persistent 4096-bit
 * Frobenius kernel), CB200_TMA=0/1 (Horner series staged by cp.async.bulk), CB200_HOST_ROWS=<rows per copy range of
 * cb200_fuchs_batch_host>. */
size_t cb200_fuchs_workspace_bytes(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                   int shared_ode);
size_t cb200_frobenius_workspace_bytes(int L, int64_t batch);
size_t cb200_horner_workspace_bytes(int L, int64_t npts);
size_t cb200_ew_workspace_bytes(int L, int64_t n);

/* ---- acb_ode_solve_fuchs for a batch (reference src/fuchs_solver.c:96-145) ----
 * ode : (order+1)*(degree+1) entries, row-major as acb_ode_coeff (reference src/acb_ode.h:21-22),
 *       S = 2*batch, or S = 2 when shared_ode != 0 (one operator for the whole batch).
 * res : n+1 entries, S = 2*batch.  On entry, entries 0 .. -valuation-1 hold the initial values
 *       (reference doc/source/cascade.rst:14-16); entries -valuation .. n are written.
 * valuation must be the value acb_ode_valuation (reference src/acb_ode.c:159-181) returns.
 * SINGLE-VALUATION CONTRACT: a launch runs with ONE valuation.  With shared_ode == 0 every operator of the batch
 * must have that same valuation (exact-zero coefficients can make valuations differ between instances: group the
 * instances by valuation and call once per group; cascade_b200.api.batch_valuation checks and rejects a mixed
 * batch).  The same holds for the Frobenius and indicial-polynomial entry points below. */
int cb200_fuchs_batch_dev(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                          const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                          uint32_t *res_mant, cb_meta *res_meta, void *workspace,
                          size_t workspace_bytes, void *stream);
/* The same solve restricted to coefficients n_begin .. n (shared operator only): res[0 .. n_begin-1] must
 * already hold the earlier coefficients (every coefficient is formed from the previous order+degree ones read
 * back from res, so the recurrence resumes anywhere).  Lets a caller solve in coefficient ranges and copy the
 * finished rows of res -- contiguous in this layout -- to the host while the next range is being computed
 * (bench.py's end-to-end leg).  Ranges give bit-identical results to one full call. */
int cb200_fuchs_batch_range_dev(int L, int order, int degree, int valuation, int64_t n_begin, int64_t n, int64_t batch,
                                const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                uint32_t *res_mant, cb_meta *res_meta, void *workspace, size_t workspace_bytes,
                                void *stream);
/* Host buffers in and out -- the call a host program makes (reference interface: acb_ode_solve_fuchs(res, ODE, deg,
 * bits), src/cascade.h:13, for a batch).  res: the caller's FULL result arrays; rows 0 .. -valuation-1 are read, rows
 * -valuation .. n are written.  Only the operator and the initial rows are uploaded; the solve runs in coefficient
 * ranges and each finished range is copied into res on a second stream while the next range is computed.  Page-locked
 * result arrays (cb200_host_alloc below, cudaHostAlloc, pinned torch tensors) make those copies run at PCIe speed;
 * pageable arrays are accepted (the driver stages them).  Device buffers, two streams and the events are kept between
 * calls; cb200_host_cache_release() frees them. */
int cb200_fuchs_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                           const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                           uint32_t *res_mant, cb_meta *res_meta, int device);
/* page-locked host memory without CUDA headers (NULL on failure), and the release of the cached device buffers */
void *cb200_host_alloc(size_t bytes);
void cb200_host_free(void *p);
void cb200_host_cache_release(void);

/* ---- acb_ode_solve_fuchs for operators with exact real INTEGER coefficients (acb_ode_legendre,
 * reference src/examples.c:3-11): ode_int[(order+1)*(degree+1)][batch] int64 (entry-major; one column
 * when shared_ode).  Same balls as cb200_fuchs_batch_*, computed with O(L) work per coefficient.
 * Every multiplier the reference forms at src/fuchs_solver.c:121-135 must stay below 2^62 in magnitude:
 * (order+1) * max|coefficient| * (n+1)^order < 2^62.  The _host entry point scans ode_int and returns CB200_ERANGE
 * otherwise; the _dev entry point cannot read device memory without synchronising, so there the KERNEL checks each
 * instance's coefficients and writes indeterminate balls (CB_FLAG_NAN) for an instance that violates the bound --
 * never a silently wrapped result. */
int cb200_fuchs_intop_batch_dev(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                const int64_t *ode_int, int shared_ode, uint32_t *res_mant,
                                cb_meta *res_meta, void *stream);
int cb200_fuchs_intop_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                 const int64_t *ode_int, int shared_ode, uint32_t *res_mant,
                                 cb_meta *res_meta, int device);

/* ---- _acb_ode_solve_frobenius, homogeneous case, M = 1 (reference src/frobenius_solver.c:72-121)
 * rho : 1 entry, S = 2*batch or 2 (shared_rho).  res: n+1 entries, all written (g_0 = 1). */
int cb200_frobenius1_batch_dev(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                               const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                               const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                               uint32_t *res_mant, cb_meta *res_meta, void *workspace,
                               size_t workspace_bytes, void *stream);
int cb200_frobenius1_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                                uint32_t *res_mant, cb_meta *res_meta, int device);

/* ---- _acb_ode_solve_frobenius with a right-hand side series (reference src/frobenius_solver.c:86-96,
 * :100): rhs = rhs_len entries, S = 2*batch or 2 (shared_rhs).  An instance whose rhs is the zero
 * polynomial takes the homogeneous branch (g_0 = 1); otherwise rho is replaced by rho - valuation and
 * g_0 = rhs_0 / f_0(rho), as the reference does.  rhs must not alias res. */
int cb200_frobenius1_rhs_batch_dev(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                   const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                   const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                                   const uint32_t *rhs_mant, const cb_meta *rhs_meta, int64_t rhs_len,
                                   int shared_rhs, uint32_t *res_mant, cb_meta *res_meta, void *workspace,
                                   size_t workspace_bytes, void *stream);
int cb200_frobenius1_rhs_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                    const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                    const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                                    const uint32_t *rhs_mant, const cb_meta *rhs_meta, int64_t rhs_len,
                                    int shared_rhs, uint32_t *res_mant, cb_meta *res_meta, int device);

/* ---- acb_ode_solve_frobenius for M = mul + alpha > 1 (the logarithmic case; reference
 * src/frobenius_solver.c:131-205 with indicial_polynomial :4-39 and _acb_ode_solution_update /
 * _extend / _normalize, src/acb_ode_solution.c:60-123).  The recurrence is carried as polynomials
 * in rho per instance, in device workspace.
 * rho  : 1 entry, S = 2*batch or 2 (shared_rho).
 * gens : M*(n+1) entries, S = 2*batch, all written: gens[i][j] (sol->gens + i, coefficient j) is
 *        entry i*(n+1) + j.  1 <= M <= 16, 1 <= mul <= M, 1 <= degree <= 8. */
size_t cb200_frobenius_general_workspace_bytes(int L, int order, int degree, int M, int64_t n, int64_t batch);
int cb200_frobenius_general_batch_dev(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                      const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                      const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                                      int mul, int M, uint32_t *gens_mant, cb_meta *gens_meta, void *workspace,
                                      size_t workspace_bytes, void *stream);
int cb200_frobenius_general_batch_host(int L, int order, int degree, int valuation, int64_t n, int64_t batch,
                                       const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode,
                                       const uint32_t *rho_mant, const cb_meta *rho_meta, int shared_rho,
                                       int mul, int M, uint32_t *gens_mant, cb_meta *gens_meta, int device);

/* ---- the solution bookkeeping steps on their own, for a batch of solutions
 * (reference src/acb_ode.h:58-60).  Polynomials cross this interface as fixed-capacity coefficient
 * arrays; their length is implied (trailing exact zeros do not count), as acb_poly normalises.
 * which: CB200_SOL_EXTEND  _acb_ode_solution_extend(sol, nu, g_nu)   src/acb_ode_solution.c:97-109
 *        CB200_SOL_UPDATE  _acb_ode_solution_update(sol, f)          src/acb_ode_solution.c:60-95
 *        CB200_SOL_NORMALIZE _acb_ode_solution_normalize(sol)        src/acb_ode_solution.c:111-123
 * gens : M*cap entries (gens[i][j] = entry i*cap + j), S = 2*batch, in/out.
 * f    : fcap entries, S = 2*batch: the polynomial g_nu / f in rho, CONSUMED (comes back zero) as in
 *        the reference; ignored for NORMALIZE.  rho: 1 entry, S = 2*batch.  0 <= M <= 16. */
#define CB200_SOL_EXTEND 0
#define CB200_SOL_UPDATE 1
#define CB200_SOL_NORMALIZE 2
size_t cb200_solution_op_workspace_bytes(int L, int M, int64_t batch);
int cb200_solution_op_dev(int which, int L, int64_t batch, int mul, int M, int64_t cap, int64_t nu,
                          const uint32_t *rho_mant, const cb_meta *rho_meta, uint32_t *gens_mant,
                          cb_meta *gens_meta, uint32_t *f_mant, cb_meta *f_meta, int64_t fcap, void *workspace,
                          size_t workspace_bytes, void *stream);
int cb200_solution_op_host(int which, int L, int64_t batch, int mul, int M, int64_t cap, int64_t nu,
                           const uint32_t *rho_mant, const cb_meta *rho_meta, uint32_t *gens_mant,
                           cb_meta *gens_meta, uint32_t *f_mant, cb_meta *f_meta, int64_t fcap, int device);

/* ---- indicial_polynomial(result, ODE, nu, shift) (reference src/frobenius_solver.c:4-39) for a
 * batch of operators: out = order+2 entries, S = 2*batch (coefficients of rho^0 .. rho^order; the
 * last entry is scratch and comes back zero). */
int cb200_indicial_polynomial_dev(int L, int order, int degree, int valuation, int64_t batch,
                                  const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode, int64_t nu,
                                  int64_t shift, uint32_t *out_mant, cb_meta *out_meta, void *workspace,
                                  size_t workspace_bytes, void *stream);
int cb200_indicial_polynomial_host(int L, int order, int degree, int valuation, int64_t batch,
                                   const uint32_t *ode_mant, const cb_meta *ode_meta, int shared_ode, int64_t nu,
                                   int64_t shift, uint32_t *out_mant, cb_meta *out_meta, int device);

/* ---- acb_ode_solution_evaluate (reference src/acb_ode_solution.c:24-58) of ONE solution at npts points:
 *   out = a^rho * sum_i C(M,i)-weighted gens[i](a) log(a)^(M-1-i)   (the weights as the reference codes them)
 * with acb_log and acb_pow computed on the device (algorithms: oracle/cbo_trans.c, cascade_b200/csrc/cb_trans.cuh;
 * Arb's own log/exp balls are implementation defined, any rigorous enclosure is admissible).
 * gens : M*cap entries, S = 2 (gens[i][j] = entry i*cap + j; lengths implied by trailing exact zeros).
 * rho  : 1 entry, S = 2.  pow_mode: 0 rho is the exact zero (a^rho = 1); 1 rho is the exact integer pow_e,
 *        0 <= pow_e < 2^31 (binary exponentiation); 2 general (exp(rho log a)).  The caller inspects rho.
 * pts  : 1 entry, S = 2*npts.  out: 1 entry, S = 2*npts.  An exactly zero point gives an indeterminate ball
 * (reference :29-33); so does a point whose ball contains zero when a logarithm is needed. */
size_t cb200_evaluate_workspace_bytes(int L, int64_t npts);
int cb200_solution_evaluate_dev(int L, int M, int64_t cap, int64_t npts, const uint32_t *gens_mant,
                                const cb_meta *gens_meta, const uint32_t *rho_mant, const cb_meta *rho_meta,
                                int pow_mode, uint64_t pow_e, const uint32_t *pts_mant, const cb_meta *pts_meta,
                                uint32_t *out_mant, cb_meta *out_meta, void *workspace, size_t workspace_bytes,
                                void *stream);
int cb200_solution_evaluate_host(int L, int M, int64_t cap, int64_t npts, const uint32_t *gens_mant,
                                 const cb_meta *gens_meta, const uint32_t *rho_mant, const cb_meta *rho_meta,
                                 int pow_mode, uint64_t pow_e, const uint32_t *pts_mant, const cb_meta *pts_meta,
                                 uint32_t *out_mant, cb_meta *out_meta, int device);
/* elementwise acb_exp (which = 0) / acb_log (which = 1) over n complex balls (1 entry, S = 2*n each) */
#define CB200_TRANS_EXP 0
#define CB200_TRANS_LOG 1
int cb200_trans_dev(int which, int L, int64_t n, uint32_t *z_mant, cb_meta *z_meta, const uint32_t *x_mant,
                    const cb_meta *x_meta, void *workspace, size_t workspace_bytes, void *stream);
int cb200_trans_host(int which, int L, int64_t n, uint32_t *z_mant, cb_meta *z_meta, const uint32_t *x_mant,
                     const cb_meta *x_meta, int device);

/* ---- acb_ode_apply / acb_ode_solves (reference src/acb_ode.c:185-240) for a batch: the reference's own
 * acceptance test of both solvers, on the device, so a sweep can verify itself without copying series back.
 * in  : len_in entries, S = 2*batch (the series).  out: out_len entries, S = 2*batch: the first out_len
 *       coefficients of sum_i p_i * in^(i).  solved (may be NULL): int32[batch], 1 where each of the first `deg`
 *       coefficients of out is finite and contains zero (acb_ode_solves(ODE, res, deg)). */
size_t cb200_ode_apply_workspace_bytes(int L, int64_t len_in, int64_t batch);
int cb200_ode_apply_dev(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant, const cb_meta *ode_meta,
                        int shared_ode, const uint32_t *in_mant, const cb_meta *in_meta, int64_t len_in,
                        uint32_t *out_mant, cb_meta *out_meta, int64_t out_len, int64_t deg, int32_t *solved,
                        void *workspace, size_t workspace_bytes, void *stream);
int cb200_ode_apply_host(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant, const cb_meta *ode_meta,
                         int shared_ode, const uint32_t *in_mant, const cb_meta *in_meta, int64_t len_in,
                         uint32_t *out_mant, cb_meta *out_meta, int64_t out_len, int64_t deg, int32_t *solved,
                         int device);

/* ---- Horner evaluation batched over points (acb_poly_evaluate inside
 * acb_ode_solution_evaluate, reference src/acb_ode_solution.c:40,45; and the first `nder`
 * coefficients of acb_poly_taylor_shift, reference src/fuchs_solver.c:159).
 * f : len entries, S = 2 (shared_f) or 2*npts.  pts: 1 entry, S = 2*npts.
 * out : nder entries, S = 2*npts; out[k] = f^(k)(pt)/k!.  1 <= nder <= 4. */
int cb200_horner_batch_dev(int L, int64_t len, int64_t npts, int nder, const uint32_t *f_mant,
                           const cb_meta *f_meta, int shared_f, const uint32_t *pts_mant,
                           const cb_meta *pts_meta, uint32_t *out_mant, cb_meta *out_meta,
                           void *workspace, size_t workspace_bytes, void *stream);
int cb200_horner_batch_host(int L, int64_t len, int64_t npts, int nder, const uint32_t *f_mant,
                            const cb_meta *f_meta, int shared_f, const uint32_t *pts_mant,
                            const cb_meta *pts_meta, uint32_t *out_mant, cb_meta *out_meta,
                            int device);

/* ---- radius_of_convergence (reference src/fuchs_solver.c:3-54) for a batch of operators, RIGOROUS on the device
 * ball type: n Graeffe iterations on the reversed leading polynomial (rescaled by a power of two after every step; as
 * many of the n requested as the 32-bit exponent range allows -- n_done[i] <= n of them, -1 when the operator has no
 * singularity away from zero and the ball is indeterminate, :17-22), the sum form of Fujiwara's root bound, its
 * 2^n_done-th root inverted; the sharpness of the bound gives the error bar, so the ball CONTAINS the distance of the
 * nearest nonzero root of the leading coefficient.  out : 1 entry, S = 2*batch (imaginary parts exact zeros). */
size_t cb200_radius_workspace_bytes(int L, int degree, int64_t batch);
int cb200_radius_of_convergence_dev(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant,
                                    const cb_meta *ode_meta, int shared_ode, int64_t n, uint32_t *out_mant,
                                    cb_meta *out_meta, int32_t *n_done, void *workspace, size_t workspace_bytes,
                                    void *stream);
int cb200_radius_of_convergence_host(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant,
                                     const cb_meta *ode_meta, int shared_ode, int64_t n, uint32_t *out_mant,
                                     cb_meta *out_meta, int32_t *n_done, int device);
/* ---- truncation_order (reference src/fuchs_solver.c:56-94) for a batch of (eta, alpha) pairs: the ball
 * N = (log(1 - r) - bits log 2) / log r, r = eta / ((eta + alpha) / 2), and n_out[i] = ceil of its upper end (2*bits for
 * non-finite input, 0 when eta < alpha cannot be shown -- the reference's early returns).  eta, alpha, n: 1 entry each,
 * S = 2*batch (real parts used). */
size_t cb200_truncation_workspace_bytes(int L, int64_t batch);
int cb200_truncation_order_dev(int L, int64_t batch, const uint32_t *eta_mant, const cb_meta *eta_meta,
                               const uint32_t *alpha_mant, const cb_meta *alpha_meta, int64_t bits, uint32_t *n_mant,
                               cb_meta *n_meta, int64_t *n_out, void *workspace, size_t workspace_bytes, void *stream);
int cb200_truncation_order_host(int L, int64_t batch, const uint32_t *eta_mant, const cb_meta *eta_meta,
                                const uint32_t *alpha_mant, const cb_meta *alpha_meta, int64_t bits, uint32_t *n_mant,
                                cb_meta *n_meta, int64_t *n_out, int device);

/* ---- acb_ode_shift (reference src/acb_ode.c:124-135) for a batch: out = in with every row p_i(z) replaced by
 * p_i(z + a), by the Horner Taylor shift p[j] += p[j+1]*a, in ONE launch (one thread per operator).
 * ode : (order+1)(degree+1) entries, S = 2*batch, or S = 2 (shared_ode: one operator moved to `batch` points).
 * a   : 1 entry, S = 2*batch or 2 (shared_a).  out : S = 2*batch.  An exactly zero a leaves the copy (:128). */
size_t cb200_ode_shift_workspace_bytes(int L, int64_t batch);
int cb200_ode_shift_dev(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant, const cb_meta *ode_meta,
                        int shared_ode, const uint32_t *a_mant, const cb_meta *a_meta, int shared_a, uint32_t *out_mant,
                        cb_meta *out_meta, void *workspace, size_t workspace_bytes, void *stream);
int cb200_ode_shift_host(int L, int order, int degree, int64_t batch, const uint32_t *ode_mant, const cb_meta *ode_meta,
                         int shared_ode, const uint32_t *a_mant, const cb_meta *a_meta, int shared_a, uint32_t *out_mant,
                         cb_meta *out_meta, int device);

/* ---- acb_poly_taylor_shift (reference src/fuchs_solver.c:159): the COMPLETE re-expansion f(x + a) of `batch` series
 * of len coefficients, in place, in the Horner form  for i = len-2..0: for j = i..len-2: f[j] += f[j+1]*a
 * (one CTA per series, each sweep i as two parallel phases; O(len^2) ball operations per series).
 * f : len entries, S = 2*batch.  a : 1 entry, S = 2*batch or 2 (shared_a). */
size_t cb200_taylor_shift_workspace_bytes(int L, int64_t len, int64_t batch);
int cb200_taylor_shift_dev(int L, int64_t len, int64_t batch, uint32_t *f_mant, cb_meta *f_meta, const uint32_t *a_mant,
                           const cb_meta *a_meta, int shared_a, void *workspace, size_t workspace_bytes, void *stream);
int cb200_taylor_shift_host(int L, int64_t len, int64_t batch, uint32_t *f_mant, cb_meta *f_meta, const uint32_t *a_mant,
                            const cb_meta *a_meta, int shared_a, int device);

/* ---- analytic_continuation (reference src/fuchs_solver.c:147-163) for a batch of solutions of ONE
 * operator along ONE path, device resident: per corner acb_ode_shift (src/acb_ode.c:124-135, one launch),
 * acb_ode_solve_fuchs with n terms, and the re-expansion at the next corner.  Only the first `order`
 * Taylor coefficients survive a step (the next solve overwrites the rest), so between corners the re-expansion is
 * a Horner evaluation with `order` normalised derivatives instead of a full acb_poly_taylor_shift.
 * ode : (order+1)(degree+1) entries, S = 2.  path : path_len entries (corners), S = 2.
 * y   : order entries, S = 2*batch: Taylor coefficients at path[0] in, at path[path_len-1] out.
 * As in the reference (src/fuchs_solver.c:105) every solve runs with the valuation of the operator SHIFTED to that
 * corner: the meta records of the shifted operator (a few hundred bytes) are read back once per corner -- the stream
 * is synchronised there, the one exception to "_dev functions do not synchronise".  An ordinary point has valuation
 * -order; at a corner where the shifted operator has a smaller |valuation| only the first -valuation entries of y
 * are read (the reference's res[0 .. -v-1]); a corner with valuation >= 0 returns CB200_EVALUATION (undefined in the
 * reference).  An exactly zero corner leaves the operator unshifted (src/acb_ode.c:128).
 * corner_is_zero is ignored (kept for ABI stability; may be NULL).  1 <= order <= 4.
 * The _series variants also return what the reference leaves in `res` after the last corner: res (n+1 entries,
 * S = 2*batch) = the complete series re-expanded at path[path_len-1] (cb200_taylor_shift); path_len >= 2. */
size_t cb200_continuation_workspace_bytes(int L, int order, int degree, int64_t n, int64_t batch);
size_t cb200_continuation_series_workspace_bytes(int L, int order, int degree, int64_t n, int64_t batch);
int cb200_continuation_dev(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                           const uint32_t *ode_mant, const cb_meta *ode_meta, const uint32_t *path_mant,
                           const cb_meta *path_meta, const uint8_t *corner_is_zero, uint32_t *y_mant,
                           cb_meta *y_meta, void *workspace, size_t workspace_bytes, void *stream);
int cb200_continuation_host(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                            const uint32_t *ode_mant, const cb_meta *ode_meta, const uint32_t *path_mant,
                            const cb_meta *path_meta, uint32_t *y_mant, cb_meta *y_meta, int device);
int cb200_continuation_series_dev(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                  const uint32_t *ode_mant, const cb_meta *ode_meta, const uint32_t *path_mant,
                                  const cb_meta *path_meta, uint32_t *y_mant, cb_meta *y_meta, uint32_t *res_mant,
                                  cb_meta *res_meta, void *workspace, size_t workspace_bytes, void *stream);
int cb200_continuation_series_host(int L, int order, int degree, int64_t n, int64_t batch, int64_t path_len,
                                   const uint32_t *ode_mant, const cb_meta *ode_meta, const uint32_t *path_mant,
                                   const cb_meta *path_meta, uint32_t *y_mant, cb_meta *y_meta, uint32_t *res_mant,
                                   cb_meta *res_meta, int device);

/* ---- elementwise ball operations over n complex balls (1 entry, S = 2*n each); k: n integers
 * for the *_SI ops (device pointer in _dev, may be NULL otherwise); y ignored for INV/_SI. */
int cb200_ew_dev(int op, int L, int64_t n, uint32_t *z_mant, cb_meta *z_meta, const uint32_t *x_mant,
                 const cb_meta *x_meta, const uint32_t *y_mant, const cb_meta *y_meta,
                 const int64_t *k, void *workspace, size_t workspace_bytes, void *stream);
int cb200_ew_host(int op, int L, int64_t n, uint32_t *z_mant, cb_meta *z_meta, const uint32_t *x_mant,
                  const cb_meta *x_meta, const uint32_t *y_mant, const cb_meta *y_meta,
                  const int64_t *k, int device);

#ifdef __cplusplus
}
#endif
#endif
