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
 *   - No allocation, no ownership transfer, no exceptions across the boundary.  The only global
 *     state is a per-device constant table (FFT twiddles) created on first use, under a mutex.
 *   - Return value: DAM_OK (0) or a negative dam_status; dam_last_error() returns a
 *     thread-local human-readable message for the last failing call on this thread.
 *   - Strides are in ELEMENTS of the tensor's dtype, sizes in elements, never bytes.
 */
#ifndef DAM_B200_H
#define DAM_B200_H

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
 *                channels > 1 are averaged (np.mean over the channel axis) in the source
 *                precision (float for I16/F32, double for F64) before framing.
 *   input_scale  multiplies every loaded sample (use 1/32768 for PCM16, 1 otherwise).
 *   row_gain     optional [n_rows] float64 gains applied after the down-mix (NULL = none).
 *   window       [n_fft] float32 analysis window (torch.hann_window(n_fft) in the reference;
 *                promoted to the compute dtype exactly as torch.stft does).
 *   n_fft, hop   window size and hop; frames are centred (reflect pad n_fft/2), so
 *                T = 1 + n_samples / hop and n_samples must exceed n_fft / 2.
 *   amin, multiplier   dB = multiplier * log10(max(|X|, amin))   (1e-5 and 20 in the reference).
 *   compute_dtype  DAM_F32 or DAM_F64 arithmetic for the FFT (the reference's dtype follows its
 *                input: float32 on the inference path, float64 on the training path).
 *   out          [n_rows, n_fft/2+1, T] contiguous, T innermost, dtype out_dtype (F32 or F64;
 *                F64 output requires compute_dtype F64).
 *   stats        optional device double[3]: sum, sum of squares and count of the dB values
 *                written are ATOMICALLY ADDED (caller zeroes it); NULL = skip.
 */
int dam_stft_db(const void* audio, int audio_dtype, int channels,
                int64_t n_rows, int64_t n_samples,
                int64_t row_stride, int64_t chan_stride, int64_t samp_stride,
                double input_scale, const double* row_gain,
                const float* window, int n_fft, int hop,
                double amin, double multiplier, int compute_dtype,
                void* out, int out_dtype, double* stats, void* stream);

/* Number of frames dam_stft_db produces for n_samples (1 + n_samples / hop). */
int64_t dam_stft_num_frames(int64_t n_samples, int hop);

#ifdef __cplusplus
}
#endif
#endif /* DAM_B200_H */
