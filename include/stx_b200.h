/*
 * libstx_b200 — C ABI of the B200-native (sm_100a) log-mel front end and cosine scorer.
 *
 * This header is the drop-in boundary for ONE hot path of
 * yuriyvnv/speech_transcript_embeddings (reference = R/, third-party transformers = TF/):
 *
 *   raw 16 kHz float32 PCM -> input_features + attention mask     (R/processor.py:79-126,
 *                                                                   R/training/trainer_unfreeze.py:855-866)
 *   cosine similarity of L2-normalised audio/text embeddings       (R/processor.py:148-159,
 *                                                                   R/inference.py:121, R/cv_inference.py:105)
 *
 * The reference has no FFI: the path sits behind a Python object protocol (the HuggingFace
 * feature-extractor __call__ and AudioTextProcessor's methods).  The Python host in
 * speech_transcript_embeddings_b200/ mirrors that protocol and binds these symbols with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; every `d_*` pointer is a DEVICE pointer, `h_*` a HOST pointer;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) unless the name ends in `_host`, which synchronises before returning;
 *   - return 0 on success, <0 on argument errors (STX_E*), >0 = a cudaError_t;
 *     stx_last_error() returns a thread-local message for the last non-zero return;
 *   - the caller allocates outputs and workspaces (ownership stays with the caller's allocator).
 */
#ifndef STX_B200_H_
#define STX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STX_ABI_VERSION 1

#define STX_EINVAL   (-1)  /* bad argument                                  */
#define STX_ENOSPACE (-2)  /* workspace too small                           */
#define STX_EDEVICE  (-3)  /* not an sm_100 device / CUDA runtime unusable  */

/* recipe K constants (SeamlessM4TFeatureExtractor, TF/models/seamless_m4t/feature_extraction_seamless_m4t.py:60-87) */
#define STX_K_FRAME   400
#define STX_K_HOP     160
#define STX_K_NFFT    512
#define STX_K_NMEL    80
#define STX_K_STRIDE  2
/* recipe W constants (WhisperFeatureExtractor, TF/models/whisper/feature_extraction_whisper.py:69-103) */
#define STX_W_NFFT    400
#define STX_W_HOP     160
#define STX_W_NMEL    80

int         stx_abi_version(void);
const char* stx_last_error(void);

/* Number of CUDA kernels this library has launched since load (all entry points); bench.py's
 * `gpu_launches` is the difference of this counter across the timed region. */
uint64_t    stx_kernel_launch_count(void);

/* Per-launch timing for bench.py's roofline leg.  While enabled, every kernel launch of the library
 * is bracketed by CUDA events on its own stream.  stx_profile_collect synchronises on those events,
 * writes up to `cap` records (kernel name, 32 bytes each, NUL-terminated; elapsed milliseconds), clears
 * the log and returns the number of records written. */
int         stx_profile_enable(int on);
int         stx_profile_collect(char* h_names, float* h_ms, int cap);

/* Copies a host-side float64 table used by the kernels, for inspection by tests.
 *   "k_window"  [400]      Povey window                      (TF/audio_utils.py:593, 601-602)
 *   "k_mel"     [257*80]   Kaldi mel filters, row-major      (TF/audio_utils.py:516-530)
 *   "w_window"  [400]      periodic Hann                     (TF/models/whisper/...:141)
 *   "w_mel"     [201*80]   Slaney mel filters, row-major     (TF/models/whisper/...:95-103)
 * Returns the number of doubles written (<= cap), or STX_EINVAL. */
int64_t     stx_get_table(const char* name, double* h_out, int64_t cap);

/* Host-side packing for the reference-signature call: the reference hands the extractor a list of ordinary (pageable)
 * float32 arrays (R/processor.py:88-105, R/training/trainer_unfreeze.py:856-860); before the H2D copy can run at PCIe speed
 * they have to be gathered into one pinned staging buffer.  Copies segment i (n_bytes[i] bytes at h_src[i]) to
 * h_dst_base + dst_byte_offsets[i] with up to `threads` native threads of a persistent in-library pool (non-temporal stores,
 * work split by bytes, not by segment).  Synchronous: the bytes are in place when it returns.  No CUDA call is made. */
int         stx_host_pack(const void* const* h_src, const int64_t* n_bytes, void* h_dst_base, const int64_t* dst_byte_offsets,
                          int count, int threads);
/* The same copy as a pipelined job: segments [chunk_starts[c], chunk_starts[c + 1]) form chunk c (n_chunks + 1 entries); the
 * chunks are packed in order on a native thread.  _begin returns a job handle (NULL on error) at once; _wait blocks until
 * chunk `chunk` is in place (chunk < 0: all of them); _end joins the job and frees it (the sources and the destination must
 * stay valid until then).  The caller overlaps the H2D copy of chunk c with the packing of chunk c + 1. */
void*       stx_host_pack_begin(const void* const* h_src, const int64_t* n_bytes, void* h_dst_base,
                                const int64_t* dst_byte_offsets, int count, const int32_t* chunk_starts, int n_chunks,
                                int threads);
int         stx_host_pack_wait(void* h_job, int chunk);
int         stx_host_pack_end(void* h_job);

/* ---------------------------------------------------------------------------------------------
 * Recipe K: Kaldi-style fbank + per-clip per-bin CMVN + pad + stride-2 stacking + mask.
 * Replaces SeamlessM4TFeatureExtractor.__call__ (TF/models/seamless_m4t/
 * feature_extraction_seamless_m4t.py:141-302) as called at R/processor.py:101-105 and
 * R/training/trainer_unfreeze.py:856-860.
 *
 *   d_pcm      packed float32 PCM of all clips (NOT pre-scaled by 2^15)
 *   d_offsets  [B] start sample of each clip inside d_pcm (int64); multiples of 4 recommended
 *   d_lengths  [B] samples per clip (int32).  T_b = 1 + (len-400)/160 frames (0 if len < 400)
 *   max_length host copy of max_b len (sizes the grid; no device->host sync is ever made).  Passed NEGATED it is a promise
 *              that EVERY clip has exactly -max_length samples (e.g. a batch cut to equal windows): the library then skips
 *              the small device-side pass that compacts the work items of ragged batches.  A wrong promise costs load
 *              balance, never correctness
 *   d_peak     NULL, or [B] float32 divisors: sample = pcm / d_peak[b] in float32 before
 *              anything else (the reference's peak-normalise, R/processor.py:91-92)
 *   T_pad      EVEN number of raw frames per clip in the padded output (TF .pad with
 *              pad_to_multiple_of=2); frames >= T_pad are used for statistics but not stored
 *   normalize  1 = per-bin (x-mean)/sqrt(var_ddof1+1e-7) (…seamless_m4t.py:257-262); 0 = raw log-mel
 *   d_out      [B, T_pad/2, 160] float32 : out[b, j, 0:80] = frame 2j, out[b, j, 80:160] = frame 2j+1,
 *              rows past the clip filled with padding_value (…seamless_m4t.py:281-300)
 *   d_mask     NULL or [B, T_pad/2] int32: 1 iff frame 2j+1 < T_b (…seamless_m4t.py:292-293)
 *   d_ws       workspace of at least stx_fbank_k_workspace() bytes
 * ------------------------------------------------------------------------------------------- */
int stx_fbank_k_workspace(int B, int max_length, size_t* bytes);
int stx_fbank_k(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B,
                int max_length, const float* d_peak, int T_pad, float padding_value, int normalize,
                float* d_out, int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream);

/* The trainer's audio collate fused into the same kernels.  Replaces the per-item extractor call of
 * CommonVoiceDataset.__getitem__ (R/training/trainer_unfreeze.py:855-866: one clip per call, no peak-normalise, no
 * trim) followed by custom_collate_fn's audio padding (R/training/trainer_unfreeze.py:898-908):
 *   d_out   [B, T_pad/2, 160] float32: the clip's stacked frames as the single-clip call returns them (the second half
 *           of the last stacked frame of an odd-T clip is padding_value), ZERO rows after them
 *   d_mask  [B, T_pad/2] int64: 1 for every stacked frame of the clip (ceil(T_b / 2) of them), else 0
 * T_pad = 2 * max_b ceil(T_b / 2).  Workspace as for stx_fbank_k. */
int stx_fbank_k_collate(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B,
                        int max_length, int T_pad, float padding_value, float* d_out, int64_t* d_mask,
                        void* d_ws, size_t ws_bytes, void* stream);

/* Recipe K fused with the speech encoder's input stage: raw PCM -> hidden states of
 * Wav2Vec2BertFeatureProjection = Linear(LayerNorm(input_features)) (TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py:118-130,
 * called at :1016), [B, T_pad/2, out_dim] float32, without writing and re-reading the normalised features: the CMVN of
 * stx_fbank_k, the padding, the LayerNorm over the 160 stacked features and the TF32 hi / lo split of the tensor-core
 * contraction are one pass over the raw log-mel.  d_features (optional) receives exactly what stx_fbank_k would return,
 * d_mask (optional) its attention mask.  Other arguments as for stx_fbank_k and stx_feature_projection. */
int stx_fbank_k_projection_workspace(int B, int max_length, int T_pad, int out_dim, int want_features, size_t* bytes);
int stx_fbank_k_projection(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B, int max_length,
                           const float* d_peak, int T_pad, float padding_value, const float* d_ln_weight,
                           const float* d_ln_bias, float eps, const float* d_weight, const float* d_bias, int out_dim,
                           float* d_hidden, float* d_features, int32_t* d_mask, void* d_ws, size_t ws_bytes, void* stream);

/* Per-clip max(1, max|x|) -> d_peak[b] (float32): the divisor of R/processor.py:91-92
 * (division only happens when max|x| > 1; dividing by exactly 1.0f is the identity). */
int stx_peak_abs(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B,
                 float* d_peak, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Recipe W: Whisper log-mel.  Replaces WhisperFeatureExtractor.__call__ +
 * _torch_extract_fbank_features (TF/models/whisper/feature_extraction_whisper.py:135-164, 189-342).
 *
 *   n_samples  clip length after pad/truncate (480000 for the stock extractor); multiple of 160
 *   d_peak     NULL, or [B] float32 divisors as in stx_fbank_k
 *   d_out      [B, 80, n_samples/160] float32
 *   d_mask     NULL or [B, n_samples/160] int32 (sample mask every 160th sample, :328-337)
 *   d_ws       workspace of at least stx_logmel_w_workspace() bytes
 * ------------------------------------------------------------------------------------------- */
int stx_logmel_w_workspace(int B, int n_samples, size_t* bytes);
int stx_logmel_w(const float* d_pcm, const int64_t* d_offsets, const int32_t* d_lengths, int B,
                 int n_samples, const float* d_peak, float* d_out, int32_t* d_mask, void* d_ws,
                 size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cosine scoring.  Replaces AudioTextProcessor.compute_similarity (R/processor.py:148-159) and
 * the F.normalize + (a*b).sum(dim=1) idiom at R/model.py:326-327, R/inference.py:121,
 * R/cv_inference.py:105, R/training/trainer_unfreeze.py:1073-1074.
 *
 *   stx_cosine_pairwise   s[i]    = <a_i/max(|a_i|,1e-12), b_i/max(|b_i|,1e-12)>          [N]
 *   stx_cosine_nxm        S[i, j] = <a_i/max(|a_i|,1e-12), b_j/max(|b_j|,1e-12)>          [N, M] (ld = M)
 * A is [N, D], B is [M, D], row-major float32.  always_normalize = 0 reproduces the reference's
 * conditional (rows are re-normalised only if some row norm of that operand deviates from 1 by
 * more than 1e-4, torch.allclose semantics); 1 normalises unconditionally.
 * d_ws: at least stx_cosine_workspace() bytes.
 * ------------------------------------------------------------------------------------------- */
int stx_cosine_workspace(int N, int M, int D, size_t* bytes);
int stx_cosine_pairwise(const float* d_a, const float* d_b, int N, int D, int always_normalize,
                        float* d_s, void* d_ws, size_t ws_bytes, void* stream);
int stx_cosine_nxm(const float* d_a, const float* d_b, int N, int M, int D, int always_normalize,
                   float* d_S, void* d_ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation-time consumers of the pairwise scores, fused into one pass over the embeddings (forward only).
 * Replaces, for one batch of [B, D] audio / clean-transcript / corrupted-transcript embeddings:
 *   F.normalize x3 + s_pos, s_neg = (aud * txt).sum(dim=1)     R/training/trainer_unfreeze.py:561-563, 1206-1207
 *   to_human_readable(s, temperature, "prob") = sigmoid(s / t)  R/training/trainer_unfreeze.py:924-939, 1215-1216
 *   AlignmentAwareInfoNCE.forward                               R/training/trainer_unfreeze.py:716-741:
 *     per_sample = CE([s_pos, s_neg] / t, target 0) * (d_align_factor[b] if given: 1 - sigmoid(mean_align) * w)
 *     loss       = mean(per_sample) + corrupt_gamma * mean(relu(s_neg))        (corrupt_gamma <= 0: no penalty)
 * All outputs float32: five [B] vectors and the scalar d_loss.
 * ------------------------------------------------------------------------------------------- */
int stx_score_pos_neg(const float* d_aud, const float* d_pos, const float* d_neg, int B, int D, float temperature,
                      float corrupt_gamma, const float* d_align_factor, float* d_s_pos, float* d_s_neg,
                      float* d_hr_pos, float* d_hr_neg, float* d_per_sample, float* d_loss, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The speech encoder's input stage, directly downstream of input_features.
 * Replaces Wav2Vec2BertFeatureProjection.forward in eval mode (TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py
 * :118-130, called at :1016): norm = LayerNorm(in_dim, eps)(x); hidden = norm @ weight.T + bias.
 *   d_x [rows, in_dim] (rows = B * T', in_dim = 160), d_ln_weight / d_ln_bias [in_dim], d_weight [out_dim, in_dim]
 *   (torch.nn.Linear layout, out_dim = 1024), d_bias [out_dim] or NULL;
 *   d_hidden [rows, out_dim]; d_norm NULL or [rows, in_dim] (the module's second return value).
 * The contraction runs on tcgen05 with split-TF32 operands (float32-grade accuracy).
 * ------------------------------------------------------------------------------------------- */
int stx_feature_projection_workspace(int rows, int in_dim, int out_dim, size_t* bytes);
int stx_feature_projection(const float* d_x, const float* d_ln_weight, const float* d_ln_bias, float eps,
                           const float* d_weight, const float* d_bias, int rows, int in_dim, int out_dim,
                           float* d_hidden, float* d_norm, void* d_ws, size_t ws_bytes, void* stream);

/* Retrieval: for every row of A the k (1..8) best-scoring rows of B under the cosine score, WITHOUT
 * materialising the N x M matrix -- the tensor-core kernel's epilogue keeps the 8 best columns of every 128-column tile in
 * registers and a merge pass picks the k best of a row.  No counterpart in the reference (its call sites score pairs,
 * R/inference.py:121, R/cv_inference.py:105); equals sorting stx_cosine_nxm's row by (score descending, column ascending).
 *   d_val [N, k] float32 scores, d_idx [N, k] int32 row indices of B (-1 where M < k).  Workspace: stx_cosine_topk_workspace. */
int stx_cosine_topk_workspace(int N, int M, int D, size_t* bytes);
int stx_cosine_topk(const float* d_a, const float* d_b, int N, int M, int D, int always_normalize, int k, float* d_val,
                    int32_t* d_idx, void* d_ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU N x M scoring: this rank's stripe S[rows of a, all M] where the text embeddings b are sharded over
 * `world` ranks of one NVLink/NVSwitch box (BASELINE.json configs[4]).  The reference has no collective; this is the
 * north_star's "all-gathered N x M cosine matrix".  The all-gather is FUSED into the kernels over peer memory:
 *   - every rank owns a symmetric buffer of stx_cosine_gather_sizes().symm_bytes bytes, mapped into all peers
 *     (h_peer_symm[r] = rank r's buffer as addressable from THIS process; torch symmetric memory provides them);
 *   - the normalise + TF32-split kernel writes this rank's text planes into its slot of EVERY peer's buffer with
 *     P2P stores over NVLink and then releases flag[rank] = epoch on every peer;
 *   - the tcgen05 GEMM visits its own columns first and, per source rank, acquires that rank's flag before the
 *     first TMA load of its columns, so the math on arrived shards overlaps the transfers still in flight.
 * `epoch` must increase by 1 per call on all ranks (start at 1; the buffer must be zeroed once after allocation and
 * before any rank's first call).  The buffer holds two sets of planes + flags used alternately by epoch parity, so
 * no cross-rank barrier is needed between calls: a rank's push of call e+2 follows its GEMM of call e+1 in stream
 * order, and that GEMM has acquired every peer's e+1 flag, which the peer published after its own GEMM of call e.
 * h_counts[r] = rows of rank r's shard (<= m_cap).
 * d_multicast: NULL, or the NVSwitch multicast address of the same symmetric buffer (one `multimem.st` then reaches
 * every peer's copy instead of `world` unicast stores).
 * d_S is [n_local, sum(h_counts)] float32.  Rows are always L2-normalised.
 * ------------------------------------------------------------------------------------------- */
int stx_cosine_gather_sizes(int n_local, int m_cap, int world, int D, size_t* ws_bytes, size_t* symm_bytes);
int stx_cosine_nxm_gathered(const float* d_a, const float* d_b, int n_local, int D, int world, int rank,
                            const int32_t* h_counts, int m_cap, void* const* h_peer_symm, void* d_multicast,
                            uint32_t epoch, float* d_S, void* d_ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Resampling to the model rate on the device: the step BEFORE the log-mel path.
 * Replaces  audio_array = librosa.resample(audio_array, orig_sr=orig_sr, target_sr=16000)  (R/processor.py:82-86)
 * with the arithmetic of librosa's res_type="polyphase" = scipy.signal.resample_poly(y, up, down) with its default
 * Kaiser(5.0) low-pass of 20 * max(up, down) + 1 taps, and librosa's fix_length to ceil(n * target_sr / orig_sr)
 * samples.  (The reference's default res_type "soxr_hq" lives in the soxr C library, which cannot be restated or
 * pinned offline: see oracle/resample.py and INTEGRATION.md.)
 *
 *   stx_resample_plan    host only: up = target_sr / gcd, down = orig_sr / gcd, taps per polyphase branch, the number of
 *                        leading outputs scipy removes, the padded filter length (any pointer may be NULL)
 *   stx_resample_filter  host only: the padded float32 filter scipy hands to upfirdn -> h_out[0 .. return value)
 *   stx_resample_poly    d_in packed float32 PCM at orig_sr (clip b at d_in_offsets[b], d_in_lengths[b] samples);
 *                        d_out packed PCM at target_sr: clip b at d_out_offsets[b] with
 *                        d_out_lengths[b] = ceil(d_in_lengths[b] * up / down) samples (the caller lays the output out;
 *                        max_out_length = max_b d_out_lengths[b]);  d_peaks (optional, [B]) receives max(1, max|y|)
 *                        per clip, the divisor of R/processor.py:91-92, reduced in the same pass.
 * ------------------------------------------------------------------------------------------- */
int stx_resample_plan(int orig_sr, int target_sr, int* up, int* down, int* taps_per_phase, int* n_pre_remove,
                      int* filter_len);
long long stx_resample_filter(int orig_sr, int target_sr, float* h_out, long long capacity);
int stx_resample_poly(const float* d_in, const int64_t* d_in_offsets, const int32_t* d_in_lengths, int B, int orig_sr,
                      int target_sr, float* d_out, const int64_t* d_out_offsets, const int32_t* d_out_lengths,
                      int max_out_length, float* d_peaks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STX_B200_H_ */
