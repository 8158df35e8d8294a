/*
 * dam_b200.h -- C ABI of the B200-native deep-audio-mixer hot path.
 *
 * The reference (apelykh/deep-audio-mixer) is pure Python and has no FFI layer of its own; each
 * entry point below replaces a group of library calls the reference makes from Python, cited as
 * file:line relative to the reference root.  The Python drop-ins in deep_audio_mixer_b200/ bind
 * these with ctypes (see INTEGRATION.md for the stub a maintainer would add to the reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.
 *   - Unless a name ends in _host, every data pointer is a DEVICE pointer owned by the caller
 *     (e.g. torch.Tensor.data_ptr()); the call is asynchronous and ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - No ownership transfer, no exceptions across the boundary.  The only device-side global state is
 *     one constant table per device (FFT twiddles), created under a mutex with one cudaMalloc +
 *     synchronous cudaMemcpy on FIRST USE of dam_stft_db.  Nothing else allocates (the K-weighting
 *     response table of a filter-coefficient set is host-side and travels in the kernel parameters).
 *     Call dam_init() beforehand if that first use would otherwise sit inside a CUDA-graph stream
 *     capture (an allocation there invalidates the capture).
 *   - Return value: DAM_OK (0) or a negative dam_status; dam_last_error() returns a
 *     thread-local human-readable message for the last failing call on this thread.
 *   - Strides are in ELEMENTS of the tensor's dtype, sizes in elements, never bytes.
 */
#ifndef DAM_B200_H
#define DAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dam_status {
    DAM_OK = 0,
    DAM_ERR_INVALID_ARGUMENT = -1,
    DAM_ERR_UNSUPPORTED = -2,   /* e.g. an n_fft the kernels are not instantiated for */
    DAM_ERR_CUDA = -3
} dam_status;

typedef enum dam_dtype {
    DAM_I16 = 0,   /* PCM16; multiplied by input_scale (1/32768 reproduces librosa.load) */
    DAM_F32 = 1,
    DAM_F64 = 2
} dam_dtype;

/* Library version "major.minor.patch". */
const char* dam_version(void);
/* Message for the last error raised on the calling thread ("" if none). */
const char* dam_last_error(void);
/* Create the constant tables of K1 on the CURRENT device now (see Conventions). */
int dam_init(void);
/* Number of kernels this library has launched in this process (all devices, all threads). */
int64_t dam_launch_count(void);
/* Name of the K1 kernel variant the last dam_stft_db call dispatched ("" before the first call). */
const char* dam_last_stft_kernel(void);

/*
 * K1: fused framing + reflect padding + window + real FFT + |.| + dB, batched over rows.
 * Replaces  MultitrackAudioDataset.compute_features   data/dataset.py:132-162
 *           (torch.stft + torch.abs + torchaudio amplitude_to_DB),
 *           the mono down-mix prologue _stereo_to_mono data/dataset.py:181-183 (channels > 1),
 *           and the gain augmentation _augment_audio   data/dataset.py:164-168 (row_gain).
 *
 *   audio        n_rows rows of `channels` x n_samples samples, dtype `audio_dtype`;
 *                element (row r, channel c, sample i) is at
 *                audio[r*row_stride + c*chan_stride + i*samp_stride]
 *                (planar [C, n]: chan_stride = n, samp_stride = 1;
 *                 interleaved [n, C]: chan_stride = 1, samp_stride = C).
 *                With n_inner > 0 rows are addressed on two levels (e.g. segment x chunk): row r
 *                starts at (r / n_inner)*outer_stride + (r % n_inner)*row_stride instead of
 *                r*row_stride; n_inner = 0 means one level.
 *                channels > 1 are averaged (np.mean over the channel axis) in the source
 *                precision (float for I16/F32, double for F64) before framing.
 *   input_scale  multiplies every loaded sample (use 1/32768 for PCM16, 1 otherwise).
 *   row_gain     optional [n_rows] float64 gains applied after the down-mix (NULL = none).
 *   window       [n_fft] float32 analysis window (torch.hann_window(n_fft) in the reference;
 *                promoted to the compute dtype exactly as torch.stft does).
 *   n_fft, hop   window size and hop; frames are centred (reflect pad n_fft/2), so
 *                T = 1 + n_samples / hop and n_samples must exceed n_fft / 2.  n_fft = 2048 (the
 *                only size the reference uses) runs the fused register-resident kernel; other
 *                powers of two in [64, 8192] run a plain shared-memory radix-2 kernel with the
 *                same semantics; anything else returns DAM_ERR_UNSUPPORTED.  1 <= hop <= n_fft.
 *   amin, multiplier   dB = multiplier * log10(max(|X|, amin))   (1e-5 and 20 in the reference).
 *   compute_dtype  DAM_F32 or DAM_F64 arithmetic for the FFT (the reference's dtype follows its
 *                input: float32 on the inference path, float64 on the training path).
 *   out          n_rows planes [n_fft/2+1, T], T innermost, dtype out_dtype (F32 or F64; F64
 *                output requires compute_dtype F64); plane r starts at out + r*out_row_stride
 *                elements (0 = densely packed, (n_fft/2+1)*T).  A stride of S*F*T lets one call
 *                per stem fill a [chunks, S, F, T] model input in place.
 *   out_bin_stride  elements between consecutive bins of a plane (0 = T, dense).  A multiple of 8
 *                (32 bytes) keeps every 8-frame output segment inside one DRAM sector: with T = 173
 *                the dense layout straddles sectors for 7 of 8 bins and the kernel runs 12 % slower;
 *                the caller then views the result as out[..., :T].
 *   stats        optional device double[3]: sum, sum of squares and count of the dB values
 *                written are ATOMICALLY ADDED (caller zeroes it); NULL = skip.
 */
int dam_stft_db(const void* audio, int audio_dtype, int channels,
                int64_t n_rows, int64_t n_samples,
                int64_t row_stride, int64_t chan_stride, int64_t samp_stride,
                int64_t n_inner, int64_t outer_stride,
                double input_scale, const double* row_gain,
                const float* window, int n_fft, int hop,
                double amin, double multiplier, int compute_dtype,
                void* out, int out_dtype, int64_t out_row_stride, int64_t out_bin_stride, double* stats,
                void* stream);

/*
 * K2: zero-order-hold gain masks applied to stems + stem sum (time domain).
 * Replaces  interpolate_mask                     inference_utils.py:12-41
 *           mixed_tracks[t] = track * mask       inference_utils.py:143
 *           np.sum(np.array(list(d.values())),0) evaluation.py:59-60,
 *                                                data/listening_test_data_preparation.py:20-21
 *           RandomModel.forward (scalar gains)   models/baselines/random_model.py:11
 *
 *   stems        n_batch songs of [n_stems, n_channels, n_samples] planar audio, dtype stems_dtype
 *                (F32 or F64); batch_stride / stem_stride / chan_stride in elements, samples
 *                contiguous.  Outputs and gains are dense with the batch as leading dimension.
 *   gains        [n_batch, n_stems, n_gains] float64 amplitude gains.
 *   interp_coef  samples held per gain = int(n_samples / n_gains) as Python computes it.
 *                mask_s[i] = gains[s][min(i / interp_coef, n_gains-1)]; with n_gains == 1 and
 *                hold_single == 0 the mask is all zeros (the reference quirk);
 *                hold_single != 0 treats a single gain as a constant gain (RandomModel).
 *   compute_dtype  F64: product and sum in float64 (float64 mask path, outputs are float64 in
 *                the reference); F32: gain cast to float32, product and sum in float32
 *                (python-float x float32-array path of RandomModel).
 *   mix_out      optional [n_batch, n_channels, n_samples], dtype out_dtype: sum over stems in
 *                stem order.
 *   stems_out    optional [n_batch, n_stems, n_channels, n_samples] contiguous, dtype out_dtype.
 *   peak, sumsq  optional device double[n_batch, n_channels]: max |mix| (atomic max) and sum mix^2
 *                (atomic add) per channel; caller zeroes them.
 */
int dam_gain_mix(const void* stems, int stems_dtype, int64_t n_batch, int64_t batch_stride,
                 int64_t n_stems, int64_t n_channels,
                 int64_t n_samples, int64_t stem_stride, int64_t chan_stride,
                 const double* gains, int64_t n_gains, int64_t interp_coef, int hold_single,
                 int compute_dtype, void* mix_out, void* stems_out, int out_dtype,
                 double* peak, double* sumsq, void* stream);

/*
 * K2 followed by the meter (evaluation.py:59-64: the stem sum is measured with pyloudnorm right after it is
 * formed): dam_gain_mix that ALSO leaves, in the workspace of the K4 call that follows, the zero-state end state of
 * the K-weighting filters over every 256-sample chunk of mix_out -- K4's first pass over the signal, computed here
 * while the mix is still in registers.  Rows of the later K4 call = the [n_batch * n_channels] rows of mix_out, in
 * that order.  Follow it with dam_kweight_segments_resume on mix_out and the same workspace; the segment sums are
 * bit-identical to dam_gain_mix + dam_kweight_segments.
 *   mix_out        required.
 *   kw_coeffs_host HOST double[10] as for dam_kweight_segments.
 *   kw_workspace   DEVICE scratch of dam_kweight_workspace_bytes(n_batch * n_channels, n_samples) bytes,
 *                  16-byte aligned.
 * Layouts the fused kernel does not take (n_samples % 4 != 0, unaligned planes) run the plain mixdown and K4's
 * own first pass -- same results, one more read of the mix.
 */
int dam_gain_mix_kweight(const void* stems, int stems_dtype, int64_t n_batch, int64_t batch_stride,
                         int64_t n_stems, int64_t n_channels,
                         int64_t n_samples, int64_t stem_stride, int64_t chan_stride,
                         const double* gains, int64_t n_gains, int64_t interp_coef, int hold_single,
                         int compute_dtype, void* mix_out, void* stems_out, int out_dtype,
                         double* peak, double* sumsq,
                         const double* kw_coeffs_host, void* kw_workspace, size_t kw_workspace_bytes,
                         void* stream);

/*
 * N3: CNN head outputs -> per-chunk amplitude gains in the layout dam_gain_mix reads.
 * Replaces  scalar_dB_to_amplitude on every head output    data/dataset_utils.py:46-50, inference_utils.py:125-131
 *   m          DEVICE float32 [n_batch * n_chunks, n_stems] (row = chunk c of song b, the model's batch order)
 *   gains_out  DEVICE float64 [n_batch, n_stems, n_chunks]: (double) powf(10, 0.5 * m), the float32 power the reference
 *              computes, widened for the float64 mask.
 */
int dam_head_gains(const float* m, int64_t n_batch, int64_t n_chunks, int64_t n_stems, double* gains_out, void* stream);

/*
 * N2: S raw stems -> n_bus bus sums (summed_track += track in stem order, in the audio dtype).
 * Replaces  sum_stems  data/medleydb_data_surgery.ipynb:305-332.
 *   stems: n_batch songs of [n_stems, n_elems] (stems stem_stride apart, songs batch_stride apart),
 *   group[n_stems] (host pointer; bus index, or -1 to drop a stem),
 *   out [n_batch, n_bus, n_elems] contiguous, same dtype (F32 or F64).
 *   dtype I16 is the PCM16 ingest form: samples are converted with x * input_scale (1/32768
 *   reproduces librosa.load exactly; N1 of SURVEY 8f) and the buses are float32; input_scale is
 *   ignored for F32 / F64.
 *   mono_out, n_channels: optional [n_batch, n_bus, n_elems / n_channels] float32 (F32 / I16 stems): the
 *   channel mean of every bus (_stereo_to_mono, data/dataset.py:181-183), n_elems being n_channels planar
 *   channels.  dam_stft_db then reads mono rows; the arithmetic (channels added in order in float32, then
 *   * 1/C or / C) is the one dam_stft_db applies itself, so the features are bit-identical.  NULL = skip.
 */
int dam_bus_sum(const void* stems, int dtype, int64_t n_batch, int64_t batch_stride,
                int64_t n_stems, int64_t n_elems, int64_t stem_stride,
                const int32_t* group_host, int n_bus, double input_scale, void* out, void* mono_out, int n_channels, void* stream);

/*
 * K3: the models' tail  masked[b] = sum_i m[b,i] * x[b,i]  (dB maps times raw head outputs) and
 * its gradients.  Replaces model_scalar_1s.py:260-275, model_scalar_2s.py:117-132,
 * model_resnet.py:120-126 (8 elementwise launches in the reference).  float32; products are
 * rounded before each add, in stem order, exactly like the reference's mul/add sequence.
 *   x [B, S, P] (P = F*T contiguous), m [B, S], out [B, P].
 * Backward: grad_m[b,i] = sum_p grad_out[b,p] * x[b,i,p] (ADDED atomically to grad_m; caller
 * zeroes it); grad_x[b,i,p] = m[b,i] * grad_out[b,p] if grad_x != NULL.
 */
int dam_feature_mix_fwd(const float* x, const float* m, int64_t B, int64_t S, int64_t P,
                        float* out, void* stream);
int dam_feature_mix_bwd(const float* x, const float* m, const float* grad_out,
                        int64_t B, int64_t S, int64_t P,
                        float* grad_m, float* grad_x, void* stream);

/*
 * K6: DummyModel.forward  y = 20*log10(sum_i 10**(0.5*x_i))  (models/baselines/dummy_model.py:23-34)
 *   x [B, S, P] float32 -> out [B, P] float32.
 */
int dam_db_sum(const float* x, int64_t B, int64_t S, int64_t P, float* out, void* stream);

/*
 * K5: per-row scalar scale.  Replaces pyln.normalize.loudness (x * 10**((target-L)/20),
 * evaluation.py:64) and librosa.util.normalize (x / max|x|, inference.ipynb:219).
 *   x [n_rows, n] (row_stride apart), factor[n_rows] device float64; divide != 0 divides instead
 *   of multiplying.  Arithmetic in float64, rounded to out_dtype.
 */
int dam_scale_rows(const void* x, int dtype, int64_t n_rows, int64_t n, int64_t row_stride,
                   const double* factor, int divide, void* out, int out_dtype, void* stream);

/* One signal times a HOST scalar: out[i] = x[i] * gain in float64, rounded to out_dtype (the scalar-gain form of
 * K5: pyln.normalize.loudness with a known gain, evaluation.py:64; MeanLoudnessModel / RandomModel stem gains,
 * models/baselines/mean_loudness_model.py:10-20, random_model.py:10-14). */
int dam_scale(const void* x, int dtype, int64_t n, double gain, void* out, int out_dtype, void* stream);

/*
 * Row reductions used by the normalisations: absmax[r] = max |x[r,:]| and sumsq[r] = sum x[r,:]^2
 * in float64 (either pointer may be NULL).  Results are ACCUMULATED (atomic max / add); caller
 * zeroes them.
 */
int dam_row_reduce(const void* x, int dtype, int64_t n_rows, int64_t n, int64_t row_stride,
                   double* absmax, double* sumsq, void* stream);

/*
 * K4: BS.1770 K-weighting (two cascaded biquads, float64 state) + sums of squares of the
 * weighted signal over caller-defined sample segments, for n_rows signals in one call.
 * Replaces the body of pyloudnorm.Meter.integrated_loudness as the reference calls it
 * (data/dataset.py:118,126; evaluation.py:32,40,63; models/baselines/mean_loudness_model.py:7,14;
 * data/listening_test_data_preparation.py:16,23): scipy.signal.lfilter twice per channel and the
 * Python loop over 400 ms gating blocks.  Gating / LUFS are computed by the caller from the
 * small seg_sumsq array.
 *
 *   x            [n_rows] signals of n_samples; sample i of row r at x[r*row_stride + i*samp_stride].
 *   coeffs_host  HOST double[10]: b0,b1,b2,a1,a2 of stage 1 (high shelf) then stage 2 (high pass),
 *                already divided by a0 (what IIRfilter.generate_coefficients returns).
 *   seg_bounds   DEVICE int64[n_seg+1], ascending sample indices in [0, n_samples]; segment k is
 *                [seg_bounds[k], seg_bounds[k+1]).  A gating block is a run of segments.
 *   seg_sumsq    DEVICE double[n_rows, n_seg]; sums are ADDED (caller zeroes).
 *   workspace    DEVICE scratch of dam_kweight_workspace_bytes(n_rows, n_samples) bytes,
 *                16-byte aligned.
 */
size_t dam_kweight_workspace_bytes(int64_t n_rows, int64_t n_samples);
/* Pre-compute the (host-side) response table of one coefficient set (HOST double[10], as below); optional. */
int dam_kweight_init(const double* coeffs_host);
int dam_kweight_segments(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                         int64_t row_stride, int64_t samp_stride,
                         const double* coeffs_host, const int64_t* seg_bounds, int64_t n_seg,
                         double* seg_sumsq, void* workspace, size_t workspace_bytes, void* stream);
/* Same, for a signal whose chunk end states dam_gain_mix_kweight already left in `workspace` (x = its mix_out,
 * unit sample stride, row_stride = n_samples): starts at the scan, reads x once instead of twice.  Unlike
 * dam_kweight_segments it OVERWRITES seg_sumsq (the scan clears it), so the caller need not zero it. */
int dam_kweight_segments_resume(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                                int64_t row_stride, int64_t samp_stride,
                                const double* coeffs_host, const int64_t* seg_bounds, int64_t n_seg,
                                double* seg_sumsq, void* workspace, size_t workspace_bytes, void* stream);

/*
 * K-weighting + mean square of every 400 ms gating block in ONE call: z_out[r, j] of
 * pyloudnorm.Meter.integrated_loudness (meter.py:69-78 of pyloudnorm 0.1.0) for n_rows signals.  The block
 * geometry (num_blocks = int(round((T - T_g)/(T_g*step)) + 1), bounds int(T_g*(j*step)*rate) ...) is
 * evaluated on the host in double precision exactly as pyloudnorm does, mapped onto disjoint segments for
 * dam_kweight_segments, and the segments of each block are added on the device.
 *   rate, block_seconds, overlap   44100, 0.4 and 0.75 in the reference.
 *   z_out        DEVICE double[n_rows, dam_kweight_num_blocks(...)].
 *   workspace    DEVICE scratch of dam_kweight_blocks_workspace_bytes(...) bytes, 16-byte aligned.
 * Fails with DAM_ERR_INVALID when the audio is shorter than one block (pyloudnorm raises ValueError).
 */
int64_t dam_kweight_num_blocks(int64_t n_samples, double rate, double block_seconds, double overlap);
size_t dam_kweight_blocks_workspace_bytes(int64_t n_rows, int64_t n_samples, double rate,
                                          double block_seconds, double overlap);
int dam_kweight_blocks(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                       int64_t row_stride, int64_t samp_stride, double rate,
                       const double* coeffs_host, double block_seconds, double overlap,
                       double* z_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Gating and LUFS of BS.1770 on the device, from the segment sums of dam_kweight_segments:
 * block j of a signal is the run of segments [block_first[j], block_last[j]); absolute gate
 * -70 LUFS, relative gate -10 LU, channel weights 1,1,1,1.41,1.41 (the tail of
 * pyloudnorm.Meter.integrated_loudness).  seg_sumsq is [n_signals * n_channels, n_seg] with the
 * channels of a signal adjacent; lufs_out is DEVICE double[n_signals].  Lets batch pipelines
 * normalise (K5 reads its factors from device memory) without a host round trip.
 */
int dam_gated_loudness(const double* seg_sumsq, int64_t n_signals, int n_channels, int64_t n_seg,
                       const int64_t* block_first, const int64_t* block_last, int64_t n_blocks,
                       double inv_block_len, double* lufs_out, void* stream);
/* Same, plus the normalisation gain of evaluation.py:64 / pyln.normalize.loudness: gain_out[i] =
 * 10^((target_lufs - L_i) / 20), DEVICE double[n_signals] -- what dam_scale_rows multiplies by. */
int dam_gated_loudness_gain(const double* seg_sumsq, int64_t n_signals, int n_channels, int64_t n_seg,
                            const int64_t* block_first, const int64_t* block_last, int64_t n_blocks,
                            double inv_block_len, double target_lufs, double* lufs_out, double* gain_out,
                            void* stream);


/* Number of frames dam_stft_db produces for n_samples (1 + n_samples / hop). */
int64_t dam_stft_num_frames(int64_t n_samples, int hop);

/*
 * N1: packed little-endian PCM samples -> float samples, element for element (interleaving kept):
 * out[i] = (sample i as a signed `bits`-bit integer) * scale.  Replaces the decode half of
 * soundfile.read / librosa.load (data/dataset.py:194, data/dataset_utils.py:53-83) for the sample widths no
 * tensor dtype covers (24-bit); scale = 2^-(bits-1) reproduces soundfile.  src: n * bits/8 bytes on the device.
 */
int dam_pcm_unpack(const void* src, int bits, int64_t n, double scale, void* out, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAM_B200_H */
