/* libtikeb200 — C ABI of the B200-native ptychography hot path.
 *
 * Every entry point takes plain pointers and sizes (device pointers are what
 * a caller reads from __cuda_array_interface__['data'][0]); arrays must be
 * C-contiguous.  All work is enqueued on the given cudaStream_t (passed as
 * void*); nothing synchronises the device and no pointer is retained after
 * return.  Return value: 0 on success, negative library code or positive
 * cudaError_t otherwise; tb_last_error() gives a thread-local message.
 *
 * "Replaces" citations are file:line in AdvancedPhotonSource/tike
 * (multislice fork) under src/tike/.
 */
#ifndef TIKE_B200_H_
#define TIKE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tb_stream_t; /* cudaStream_t */

#define TB_OK 0
#define TB_ERR_INVALID (-1)
#define TB_ERR_UNSUPPORTED (-2)
#define TB_ERR_CUDA (-3)

#define TB_DATA_F32 0
#define TB_DATA_U16 1

#define TB_NOISE_GAUSSIAN 0
#define TB_NOISE_POISSON 1

#define TB_STEP_ALL_MODES 0
#define TB_STEP_DOMINANT_MODE 1

const char* tb_last_error(void);
int tb_version(void);
/* number of SMs of the current device (grid sizing for persistent kernels) */
int tb_sm_count(int* count);

/* ---- Patch operator -----------------------------------------------------
 * Replaces operators/cupy/patch.py:79-188 + convolution.cu:35-165
 * (fwd_patch / adj_patch<float2,float2,float>).
 * images (nimage, H, W) c64; positions (nimage, nscan, 2) f32;
 * patches (nimage, nscan*nrepeat | npatch, padded, padded) c64. */
int tb_patch_fwd(const void* images, void* patches, const float* positions,
                 int nimage, int height, int width, int nscan, int nrepeat,
                 int patch_width, int padded_width, tb_stream_t stream);
int tb_patch_adj(void* images, const void* patches, const float* positions,
                 int nimage, int height, int width, int nscan, int nrepeat,
                 int patch_width, int padded_width, int npatch,
                 tb_stream_t stream);

/* ---- Propagation operator ----------------------------------------------
 * Replaces operators/cupy/propagation.py:43-73 (cuFFT C2C via cache.py).
 * In-place batched 2-D FFT of (batch, n, n) c64, DC at the corner, natural
 * order in and out; result multiplied by `scale`.  n = 2^k, 16 <= n <= 2048:
 * n <= 128 one shared-memory pass, larger row/column two-pass.  Any other
 * 2 <= n <= 1024 (cuFFT takes every width): chirp-z transform on top of the
 * power-of-two path, with a transient stream-ordered scratch
 * (cudaMallocAsync / cudaFreeAsync on `stream`). */
int tb_fft2(void* x, int64_t batch, int n, int inverse, float scale,
            tb_stream_t stream);

/* ---- one batch of scan positions ---------------------------------------- */
typedef struct tb_batch {
  const void* psi;            /* (H, W) c64 object slice */
  int32_t height, width;
  const float* scan;          /* (npos, 2) f32, row then column */
  int64_t npos;
  const void* probe;          /* (M, N, N) c64 or (npos, M, N, N) */
  int32_t nmodes, probe_width;
  int32_t probe_per_position; /* 1: probe has a leading position axis */
  const void* eigen_probe;    /* (E, Me, N, N) c64 or NULL */
  int32_t neigen, eigen_modes;
  const float* eigen_weights; /* (npos, E+1, M) f32 or NULL */
  int32_t detector_width;     /* ND >= N; pad = (ND - N) / 2 */
  float fwd_scale, inv_scale; /* FFT normalisation ('ortho': 1/ND both) */
} tb_batch;

/* Ptycho.fwd / _compute_intensity (operators/cupy/ptycho.py:114-204,
 * ptycho/ptycho.py:95-124).  farplane (npos, M, ND, ND) c64 and/or
 * intensity (npos, ND, ND) f32 = sum_m |farplane|^2; either may be NULL. */
int tb_ptycho_fwd(const tb_batch* b, void* farplane, float* intensity,
                  tb_stream_t stream);

/* ---- rPIE ---------------------------------------------------------------
 * One call = rpie._get_nearplane_gradients for one batch
 * (ptycho/solvers/rpie.py:315-567) fused into one kernel: patch, probe
 * product, FFT, cost, modulus/Poisson step, inverse FFT, object and probe
 * numerators, eigen-weight increment. */
typedef struct tb_rpie_args {
  tb_batch batch;
  const void* data;           /* (npos, ND, ND) f32 or u16 */
  int32_t data_dtype;         /* TB_DATA_* */
  const uint8_t* mask;        /* (ND, ND) measured pixels, NULL = all */
  int32_t num_measured;       /* count of measured pixels (ND*ND if no mask) */
  int32_t noise_model;        /* TB_NOISE_* */
  int32_t step_mode;          /* TB_STEP_* (Poisson only) */
  float step_length_start, step_length_weight;
  float unmeasured_scaling;   /* ExitWaveOptions.unmeasured_pixels_scaling */
  int32_t accumulate_object;  /* compute psi / probe numerators */
  void* psi_numerator;        /* (H, W) c64, accumulated into */
  void* probe_numerator;      /* (M, N, N) c64, overwritten (rpie.py:349) */
  float* costs;               /* (npos,) f32 */
  float* eigen_weight_step;   /* (npos,) f32: 0.1*num/den for mode 0, or NULL */
  void* workspace;            /* tb_rpie_workspace_size() bytes */
  int64_t workspace_bytes;
} tb_rpie_args;

int64_t tb_rpie_workspace_size(const tb_rpie_args* a);
int tb_rpie_batch(const tb_rpie_args* a, tb_stream_t stream);

/* rpie._update (rpie.py:217-312) without adaptive moment:
 *   psi   += num / ((1-alpha) * precond + alpha * max(precond))
 *   probe += num / (alpha * max(probe_precond))
 * `scratch` is a device float[2] used for the max reduction. */
int tb_rpie_update_psi(void* psi, const void* numerator, const void* precond,
                       int64_t n, float alpha, float* scratch,
                       tb_stream_t stream);
int tb_rpie_update_probe(void* probe, const void* numerator,
                         const void* probe_precond, int nmodes, int64_t n2,
                         float alpha, float* scratch, tb_stream_t stream);

/* ---- preconditioners (solvers/_preconditioner.py:48-167) ----------------
 * psi_precond (H, W) c64 = scatter_s(sum_m |P_m|^2), overwritten;
 * probe_precond (N, N) c64 = sum_s |patch_s|^2, overwritten.
 * `order` (npos int32, or NULL = natural order) is the sequence in which the
 * positions are visited; it never changes the sums, only their speed.  With
 * positions sorted by (row / 16, column) and N <= 128 both run as window
 * kernels that keep the overlapping footprints of consecutive positions in
 * shared memory (csrc/precond.cu). */
int tb_precond_psi(const void* probe, int nmodes, int probe_width,
                   const float* scan, const int32_t* order, int64_t npos,
                   void* psi_precond, int height, int width,
                   float* scratch /* N*N floats */, tb_stream_t stream);
int tb_precond_probe(const void* psi, int height, int width,
                     const float* scan, const int32_t* order, int64_t npos,
                     int probe_width, void* probe_precond, tb_stream_t stream);

/* ---- lstsq_grad ----------------------------------------------------------
 * Phase 1 = lstsq._get_nearplane_gradients (lstsq.py:367-602): like rPIE but
 * the back-propagated residual chi (npos, M, N, N) is kept, the object
 * gradient has no 1/M and position-gradient sums are optional.
 * Phase 2 = _precondition_nearplane_gradients (lstsq.py:619-718): per
 * position sums A1, A4, b1, b2, A2 for the 2x2 step-length solve. */
typedef struct tb_lstsq_args {
  tb_batch batch;
  const void* data;
  int32_t data_dtype;
  const uint8_t* mask;
  int32_t num_measured;
  int32_t noise_model, step_mode;
  float step_length_start, step_length_weight;
  float unmeasured_scaling;
  int32_t recover_psi, recover_probe, recover_positions;
  void* chi;                  /* (npos, M, N, N) c64 out */
  void* object_upd_sum;       /* (H, W) c64 accumulated */
  void* probe_upd_sum;        /* (M, N, N) c64 overwritten: sum_s conj(o) chi */
  float* costs;               /* (npos,) */
  float* position_num;        /* (npos, 2) or NULL */
  float* position_den;        /* (npos, 2) or NULL */
  float gradient_taps[5];     /* Gaussian first-derivative taps, sigma 0.333 */
  void* workspace;
  int64_t workspace_bytes;
} tb_lstsq_args;

int64_t tb_lstsq_workspace_size(const tb_lstsq_args* a);
int tb_lstsq_phase1(const tb_lstsq_args* a, tb_stream_t stream);

/* out (npos, 6) f32: A1, A4, b1, b2, Re A2, Im A2 (before the 0.5*mean
 * damping).  object_update (H, W) c64 is the preconditioned object update,
 * m_probe_update (N, N) c64 is mode `mode` of the mean probe update. */
int tb_lstsq_phase2(const tb_batch* b, const void* chi,
                    const void* object_update, const void* m_probe_update,
                    int mode, float eps, float* out, tb_stream_t stream);

/* object_upd / sqrt(((1-alpha) precond)^2 + (alpha max precond)^2)
 * (lstsq.py:605-616); scratch = device float[1]. */
int tb_lstsq_precondition_object(void* out, const void* object_upd,
                                 const void* precond, int64_t n, float alpha,
                                 float* scratch, tb_stream_t stream);

/* y += a * x over n complex64 values (a real, read from device if a_dev) */
int tb_caxpy(void* y, const void* x, int64_t n, float a, const float* a_dev,
             tb_stream_t stream);

/* ---- object-sized updates and per-epoch constraints (csrc/update.cu) --------
 * Single passes over n = D*H*W complex64 values replacing chains of array
 * expressions in the reference.  The *_given_max forms take max(Re precond)
 * as a device scalar: with the object rows split over ranks the maximum is a
 * cross-rank quantity (communicators/comm.py, RowPlan); tb_max_real computes
 * the local one (out must be initialised, e.g. to 0: out = max(out, ...)). */
int tb_max_real(const void* x, int64_t n, float* out, tb_stream_t stream);
/* rpie._update object step, rpie.py:233-238 */
int tb_rpie_update_psi_given_max(void* psi, const void* numerator,
                                 const void* precond, int64_t n, float alpha,
                                 const float* precond_max, tb_stream_t stream);
/* rpie._update with use_adaptive_moment and no cost history (rpie.py:233-267,
 * opt.adam opt.py:165-213): psi += g/deno; (m, v) updated from g;
 * psi += adam(g)/deno.  v (n,) f32, m (n,) c64, both in/out.  The decays are
 * doubles so that 1 - decay is rounded to float32 once, like the scalar
 * operands of the reference's array expressions. */
int tb_rpie_update_psi_adam(void* psi, const void* numerator,
                            const void* precond, float* v, void* m, int64_t n,
                            float alpha, double vdecay, double mdecay,
                            const float* precond_max, tb_stream_t stream);
/* lstsq_grad object step with opt.momentum (lstsq.py:176-193, opt.py:67-82):
 * m = mdecay m + (1 - mdecay) beta direction; psi += m; beta on the device */
int tb_momentum_update(void* psi, const void* direction, void* m, int64_t n,
                       double mdecay, const float* beta, tb_stream_t stream);
/* lstsq._precondition_object_update, lstsq.py:605-616 */
int tb_lstsq_precondition_object_given_max(void* out, const void* object_upd,
                                           const void* precond, int64_t n,
                                           float alpha, const float* precond_max,
                                           tb_stream_t stream);
/* y[i] += numerator[i] / (Re precond[i % period] + eps); period 0 = n
 * (the one-update-per-epoch step of solvers/dm.py) */
int tb_add_quotient(void* y, const void* numerator, const void* precond,
                    int64_t n, int64_t precond_period, float eps,
                    tb_stream_t stream);
/* positivity_constraint (object.py:208-224; 0 = off) then clip_magnitude
 * (ptycho.py:257-262; clip != 0) in place */
int tb_object_pointwise_constraints(void* psi, int64_t n, float positivity,
                                    int clip, float a_max, tb_stream_t stream);
/* smoothness_constraint (object.py:227-253): 3x3 kernel, a on the neighbours,
 * 1 - 8a in the centre, edges replicated ('nearest'); out != psi */
int tb_object_smoothness(void* out, const void* psi, int nslices, int height,
                         int width, float a, tb_stream_t stream);
/* out[0] = sum |psi|^2 Re w, out[1] = sum (Re w)^2 in float64 on the device:
 * the two reductions of remove_object_ambiguity (object.py:324-335) */
int tb_weighted_norm_sums(const void* psi, const void* weight, int64_t n,
                          double* out, tb_stream_t stream);
/* y *= s (divide == 0) or y /= s, s a device float */
int tb_scale_by_device_scalar(void* y, int64_t n, const float* s, int divide,
                              tb_stream_t stream);

/* ---- lstsq_grad variable-probe (eigen probe) updates (csrc/eigen.cu) --------
 * lstsq._update_nearplane (lstsq.py:297-364, 721-761) with
 * probe.update_eigen_probe (probe.py:362-476) for one batch and one probe mode,
 * without (B, N, N) temporaries: patches, residuals and projections are
 * rebuilt per position.  eigen_probe points at eigen probe 0 of `mode`, probe k
 * lives eigen_stride complex values further; coefs (npos, ncoef) c64 holds the
 * projection coefficients <R, E_k> / <E_k, E_k> of the eigen probes already
 * updated in this batch (written by pass 2, read by both passes).
 * pass 1 (c >= 1): update (N, N) c64 += sum_s R_s (mean_px Re(conj(R_s) E_c)
 *   + w[s]) * inv_norm_weights  (zero it first; the caller divides by the
 *   batch size, normalises and refreshes E_c, probe.py:439-457);
 *   intensity_sums (npos, 2) f32 (optional, any c >= 0): sum Re(conj(o P) chi),
 *   sum |o P|^2 of the main-probe intensity coefficient (lstsq.py:721-736).
 * pass 2: n_out[s] = mean Re(chi conj(o E_c)), d_out[s] = mean |o E_c|^2 with
 *   the REFRESHED E_c (probe.py:459-473), and coefs[s, c-1] when given. */
int tb_lstsq_eigen_pass1(const tb_batch* b, const void* chi, int mode,
                         const void* m_probe_update, const void* eigen_probe,
                         int64_t eigen_stride, int c, const void* coefs,
                         int ncoef, const float* weights, int64_t weight_stride,
                         const float* inv_norm_weights /* device scalar */, void* update,
                         float* intensity_sums, tb_stream_t stream);
int tb_lstsq_eigen_pass2(const tb_batch* b, const void* chi, int mode,
                         const void* m_probe_update, const void* eigen_probe,
                         int64_t eigen_stride, int c, void* coefs, int ncoef,
                         float* n_out, float* d_out, tb_stream_t stream);

/* ---- multislice objects (D > 1), rPIE --------------------------------------
 * The slice loop of the reference fork: forward model through the slices
 * with a Fresnel-spectrum step in between
 * (operators/cupy/multislice.py:69-139, fresnelspectprop.py:52-111), the
 * per-slice gradients of rpie._get_nearplane_gradients (rpie.py:374, 441-474)
 * and the per-slice object preconditioner (_preconditioner.py:48-100).
 * Conventions: batch.psi points at (D, H, W) c64 slices; `propagator` is the
 * (N, N) c64 Fresnel spectrum kernel (DC at the corner); probe width must
 * equal the detector width; args->psi_numerator is (D, H, W), accumulated
 * into; args->probe_numerator is (D, M, N, N), overwritten; args->workspace
 * holds tb_multislice_workspace_size() bytes.  D == 1 is accepted and equals
 * the single-slice math. */
int64_t tb_multislice_workspace_size(const tb_batch* batch, int nslices);
int tb_multislice_fwd(const tb_batch* batch, int nslices, const void* propagator,
                      void* farplane, float* intensity, void* workspace,
                      int64_t workspace_bytes, tb_stream_t stream);
int tb_multislice_rpie_batch(const tb_rpie_args* args, int nslices,
                             const void* propagator, tb_stream_t stream);
/* lstsq_grad phase 1 on a multislice object as the fork runs it
 * (lstsq.py:422-530): multislice forward model, then the single-slice
 * gradients of slice 0 -- args->object_upd_sum is the (H, W) plane of slice 0,
 * args->workspace holds tb_multislice_workspace_size() bytes. */
int tb_multislice_lstsq_phase1(const tb_lstsq_args* args, int nslices,
                               const void* propagator, tb_stream_t stream);
/* psi_precond (D, H, W) c64, overwritten */
int tb_multislice_precond_psi(const tb_batch* batch, int nslices,
                              const void* propagator, void* psi_precond,
                              void* workspace, int64_t workspace_bytes,
                              tb_stream_t stream);

/* ---- host-side clustering helper ------------------------------------------
 * Growth loop of cluster.wobbly_center (cluster.py:360-376), bit-exact with
 * the reference's NumPy float32 arithmetic.  population (npoints, 2) f32 host
 * array; labels (npoints,) uint16 host array, 0xFFFF = unassigned; performs
 * `steps` assignments round-robin over the clusters. */
int tb_cluster_grow(const float* population, int64_t npoints, int ndim,
                    uint16_t* labels, int num_cluster, int64_t steps);

/* One sweep of the pairwise-swap refinement of cluster.compact
 * (cluster.py:587-626), same float64 expressions and visiting order as the
 * NumPy loop.  dist (n, num_cluster) f64; labels (n,) u16 in/out; wanted (n,)
 * i64 = argmin(dist, 1); happiness (n,) f64 in/out; order (n,) i64 =
 * argsort(happiness) at sweep start.  Host arrays.  Returns 1 if a swap
 * happened, 0 if none, < 0 on bad arguments. */
int tb_cluster_compact_sweep(const double* dist, uint16_t* labels,
                             const int64_t* wanted, double* happiness,
                             const int64_t* order, int64_t n, int num_cluster);

/* Inlier pass of the RANSAC affine fit of the position regularisation
 * (position.py:277-327): residual of `positions0 @ [[m00, m01], [m10, m11]] +
 * (t0, t1) - positions1` per position, inlier[k] = |residual|^2 <=
 * max_error_sq, in the float64 arithmetic of the NumPy expressions it replaces
 * (bit-exact).  Host arrays of n doubles; `inlier` (n bytes) may be NULL when
 * only the count is wanted. */
int tb_affine_inliers(const double* x0, const double* y0, const double* x1,
                      const double* y1, int64_t n, const double* m,
                      double t0, double t1, double max_error_sq,
                      uint8_t* inlier, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* TIKE_B200_H_ */
