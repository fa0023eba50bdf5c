/*
 * wavernn_b200.h -- C ABI of the B200-native WaveRNN (fatchord) vocoder inference engine.
 *
 * This is the drop-in boundary for the hot path named by BASELINE.json: everything below
 * vocoder.inference.infer_waveform / WaveRNN.generate of RuntimeRacer/Real-Time-Voice-Cloning.
 * Plain pointers and sizes only; no torch / pybind types.  The reference has NO C ABI for its PyTorch
 * path; its only native boundary is the pybind11 module `WaveRNNVocoder`
 * (vocoder/libwavernn/fatchord_version/src/WaveRNNVocoder.cpp:51-84: Vocoder{loadWeights,
 * setRandomSeed, melToWav}).  Each entry point cites what it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds under vocoder/inference.py.
 *
 * Threading: one engine per GPU; calls on one engine are serialised by the caller (the reference's
 * handles are not re-entrant either).  ctypes releases the GIL for the call.
 * Every function returns WRNN_OK (0) or a negative status; wrnn_last_error() gives the message.
 */
#ifndef WAVERNN_B200_H
#define WAVERNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wrnn_engine wrnn_engine;

enum {
    WRNN_OK = 0,
    WRNN_ERR_INVALID = -1,    /* bad argument (maps to ValueError / NotImplementedError in the wrapper)     */
    WRNN_ERR_NOT_LOADED = -2, /* "Please load Wave-RNN in memory before using it", vocoder/inference.py:70 */
    WRNN_ERR_CUDA = -3,       /* CUDA runtime failure                                                      */
    WRNN_ERR_TIMEOUT = -4,    /* the persistent loop's deadlock guard fired                                */
    WRNN_ERR_SHAPE = -5,      /* tensor shape mismatch (load_state_dict would raise)                       */
    WRNN_ERR_TOO_SHORT = -6   /* T <= 20 frames: the reference raises ValueError at fatchord_version.py:255 */
};

enum { WRNN_MODE_RAW = 0, WRNN_MODE_MOL = 1 };       /* hparams.mode, config/hparams.py:222              */

/* Arithmetic of the sample loop.
 *   F32  : fp32 weights resident in shared memory, fp32 FMA, fp32 state  (parity mode)
 *   F16  : fp16 weights/activations on tensor cores, fp32 accumulate + fp32 recurrent state
 *   SPARSE_F32 : block-sparse (1x4 groups, vocoder/pruner.py) fp32 variant                        */
enum { WRNN_PREC_F32 = 0, WRNN_PREC_F16 = 1, WRNN_PREC_SPARSE_F32 = 2 };

/* progress_callback(i, seq_len, b_size, gen_rate_kHz): fatchord_version.py:234-236.  Called from the
 * host thread that is inside wrnn_generate (it polls a device-written step counter).               */
typedef void (*wrnn_progress_fn)(int64_t i, int64_t seq_len, int64_t b_size, double gen_rate_khz, void* user);

/* base.init_voc_model(MODEL_TYPE_FATCHORD, device, override_hp_fatchord) -- vocoder/models/base.py:18-48.
 * Fixed topology of the named hparams (rnn_dims=fc_dims=512, compute/res_out=128, res_blocks=10,
 * upsample (5,5,8), hop 200, pad 2, feat 80); `bits` 8..10 for RAW (the reference trains 9 and 10); MOL has 30 outputs; anything else: WRNN_ERR_INVALID. */
int wrnn_create(int device, int bits, int mode, wrnn_engine** out);
int wrnn_destroy(wrnn_engine* e);
/* base.init_voc_model(MODEL_TYPE_RUNTIMERACER, ...) -- vocoder/models/base.py:82-104, the model type vocoder/inference.py:43
 * hard-codes for the reference's C++ path: four GRU-256 cells on a residual chain and five FC layers
 * (vocoder/models/runtimeracer_version.py:119-132), the fatchord front end (config/hparams.py:355-366).  Call between
 * wrnn_create and the first wrnn_set_tensor; the state_dict names are the reference's (I, rnn1..rnn4, fc1..fc5, upsample.*).
 * This topology runs WRNN_PREC_F32 only (wrnn_loop_rr_kernel).
 * WRNN_TOPO_GENEING: base.init_voc_model(MODEL_TYPE_GENEING, ...) -- base.py:57-80: one GRU-256, fc1 (288 -> 128, ReLU), fc3
 * (vocoder/models/geneing_version.py:107-113), its own front end (64 channels, 3 residual blocks, upsampling 4 x 5 x 10:
 * config/hparams.py:288-300); mode 'BITS' is WRNN_MODE_RAW (softmax over 2**bits classes), 'MOL' WRNN_MODE_MOL; the
 * beta-distribution mode is not supported.  wrnn_loop_gn_kernel, WRNN_PREC_F32 only.                                      */
enum { WRNN_TOPO_FATCHORD = 0, WRNN_TOPO_RUNTIMERACER = 1, WRNN_TOPO_GENEING = 2 };
int wrnn_set_topology(wrnn_engine* e, int topology);
/* WaveRNNVocoder::loadWeights(path) -- vocoder/libwavernn/fatchord_version/src/WaveRNNVocoder.cpp:22-31: create + load + finalize an
 * engine from a libwavernn `.bin` export (vocoder/libwavernn/convert.py; fatchord topology, usually pruned).  Mode and bits
 * follow from fc3's row count.  On failure *out is NULL and `err` (optional, err_len bytes) holds the message -- a file that
 * cannot be opened gives the reference's "Cannot open file.".                                                              */
int wrnn_create_from_bin(const char* path, int device, wrnn_engine** out, char* err, int err_len);
const char* wrnn_last_error(const wrnn_engine* e);

/* model.load_state_dict(checkpoint["model_state"]) -- vocoder/inference.py:35.  One call per entry of
 * the reference state_dict (names as in fatchord_version.py:88-118, e.g. "rnn1.weight_ih_l0");
 * float32 host data.  Integer entries ("step", "*.num_batches_tracked") are accepted and ignored
 * except "step", which is kept for get_step() (fatchord_version.py:406-407).                        */
int wrnn_set_tensor(wrnn_engine* e, const char* name, const float* data, const int64_t* shape, int ndim);
int wrnn_set_step(wrnn_engine* e, int64_t step);
int64_t wrnn_get_step(const wrnn_engine* e);

/* Finish loading: folds BatchNorm into the convolutions, builds the conditioning projections and the
 * loop weights in every precision, detects 1x4 block sparsity of a pruned checkpoint
 * (vocoder/pruner.py:60-88 leaves zeros, no mask) and uploads everything.                           */
int wrnn_finalize(wrnn_engine* e);
/* fraction of zero 1x4 groups found in [rnn1.hh, rnn2.ih, rnn2.hh, fc1, fc2, fc3] at finalize        */
double wrnn_sparsity(const wrnn_engine* e);
/* 1 when the block-sparse loop can run this checkpoint (its compressed per-CTA images fit one cluster's shared memory);
 * a pruned checkpoint that does not fit runs the dense loops instead (vocoder/pruner.py:60-88 leaves the tensors dense). */
int wrnn_sparse_available(const wrnn_engine* e);

/* Index arithmetic of WaveRNN.fold_with_overlap -- fatchord_version.py:315-326 (pure host).         */
int wrnn_fold_plan(int64_t total_len, int64_t target, int64_t overlap, int64_t* num_folds, int64_t* padded_len);

typedef struct {
    /* ---- inputs ------------------------------------------------------------------------------ */
    int32_t n_utts;             /* number of utterances in this call                                 */
    const float* const* mels;   /* n_utts pointers to (80, T[i]) row-major float32, ALREADY divided   */
                                /* by sp.max_abs_value (infer_waveform does that, inference.py:91)    */
    const int32_t* T;           /* frames per utterance                                               */
    int32_t mels_on_device;     /* 0: host pointers (copied inside the call); 1: device pointers      */
    int32_t batched;            /* generate(batched=...), fatchord_version.py:155                     */
    int32_t target, overlap;    /* fold plan; ignored when batched == 0                               */
    int32_t mu_law;             /* hp.mu_law; forced off for MOL (fatchord_version.py:156)            */
    int32_t apply_preemphasis;  /* sp.preemphasize -> de_emphasis, fatchord_version.py:249-250        */
    int32_t precision;          /* WRNN_PREC_*                                                        */
    uint64_t seed;              /* Philox key (noise contract: csrc/philox.cuh, oracle/philox.py)     */
    int32_t utt_index0;         /* Philox utterance counter of mels[0] (sharding keeps noise global)  */
    int32_t fold_begin, fold_end; /* n_utts==1 only: run folds [fold_begin, fold_end) of the          */
                                /* utterance (multi-GPU fold sharding); 0,0 = all.  With a partial    */
                                /* range only `samples` is produced (no post chain).                  */
    const float* forced;        /* optional, n_utts==1: (F,S) values fed back instead of own samples  */
    int32_t max_steps;          /* 0 = all; >0 stops the loop early (debug / teacher-forced prefixes) */
    wrnn_progress_fn progress;  /* may be NULL                                                        */
    void* progress_user;
    /* ---- outputs (any may be NULL) -------------------------------------------------------------- */
    double* wav;                /* concatenated float64 waveforms, utterance i has (T[i]-1)*200       */
    int64_t wav_capacity;       /* in samples                                                         */
    int64_t* wav_offsets;       /* n_utts+1 entries                                                   */
    int32_t wav_on_device;      /* 1: `wav` is a device pointer (HBM-resident bench leg)              */
    float* samples;             /* n_utts==1: (F,S) fed-back sample values before unfold (host)       */
    float* logits;              /* n_utts==1: (F,S,C) per-step logits (host; debug, small sizes only) */
    /* ---- timings filled by the call (milliseconds, CUDA events on the engine's stream) ----------- */
    float ms_h2d, ms_cond, ms_loop, ms_post, ms_d2h;
    int32_t n_folds, n_steps, n_launches;
    int32_t loop_kernel;        /* which loop ran the last wave: WRNN_LOOP_F32 / _TC / _RS / _SPARSE / _TC2     */
} wrnn_request;
enum { WRNN_LOOP_F32 = 0, WRNN_LOOP_TC = 1, WRNN_LOOP_RS = 2, WRNN_LOOP_SPARSE = 3, WRNN_LOOP_TC2 = 4, WRNN_LOOP_RR = 5, WRNN_LOOP_GN = 6 };

/* WaveRNN.generate(mels, batched, target, overlap, mu_law, apply_preemphasis, progress_callback)
 * -- fatchord_version.py:155-259 -- for one or many utterances, end to end on the GPU:
 * pad + MelResNet + upsample (as per-frame tables), fold, the autoregressive loop with fused
 * sampling, xfade_and_unfold, decode_mu_law, de_emphasis, truncate, fade-out.                        */
int wrnn_generate(wrnn_engine* e, wrnn_request* req);

/* Conditioning front end only (UpsampleNetwork.forward, fatchord_version.py:78-85): returns the
 * MelResNet output per frame, aux (T,128), and -- if mels_up != NULL -- the upsampled mel
 * (200*T, 80) reconstructed from the engine's interpolation tables (debug / parity).                */
int wrnn_condition(wrnn_engine* e, const float* mel, int32_t T, float* aux_frames, float* mels_up);

/* Same front end through the tensor-core path (cond_tc.cu: tcgen05 + TMA, hi/lo fp16 operand pairs, fp32
 * accumulation) that the fp16 loop uses; returns aux (T,128).                                          */
int wrnn_condition_tc(wrnn_engine* e, const float* mel, int32_t T, float* aux_frames);

/* Post chain only (fatchord_version.py:242-255) on host (F,S) float32 samples: xfade_and_unfold,
 * decode_mu_law, de_emphasis, truncation and fade-out on the GPU.  wav must hold (T-1)*200.          */
int wrnn_postprocess(wrnn_engine* e, const float* samples, int64_t num_folds, int64_t S, int32_t batched,
                     int32_t overlap, int32_t T, int32_t mu_law, int32_t apply_preemphasis, double* wav);

/* WaveRNN.xfade_and_unfold(y, target, overlap) alone -- fatchord_version.py:342-404 -- on float64 (F,S)
 * host input; `target` is recomputed from S like the reference does (Q7).  out holds
 * F*(S-overlap)+overlap doubles.  The input is not modified (the reference scales it in place).        */
int wrnn_xfade_unfold(wrnn_engine* e, const double* y, int64_t num_folds, int64_t S, int32_t overlap, double* out);

/* Measured floor of one inter-SM exchange of the persistent loop (microseconds per round):
 * ll_us: flag-in-data (8-byte value+tag words) all-gather of 512 values across the grid;
 * counter_us: fence + atomic counter grid barrier.  north_star: "per-step latency against the
 * measured grid-barrier floor".                                                                      */
int wrnn_barrier_floor(wrnn_engine* e, int32_t rounds, float* ll_us, float* counter_us);

/* tcgen05.mma issue-rate microbenchmark (M=128, K=16, one thread, `iters` groups of 4 MMAs): SM clocks spent issuing
 * and until completion.  mode 0: one commit at the end; 1: commit per group; 2: wait for every commit.            */
int wrnn_debug_umma_rate(wrnn_engine* e, int32_t N, int32_t iters, int32_t mode, int64_t* cycles_issue, int64_t* cycles_total);

/* Floor of one cluster-local exchange (DSMEM stores into every peer + hardware cluster barrier), microseconds per
 * round, for a cluster of `cluster_size` CTAs: the exchange the block-sparse loop uses.                          */
int wrnn_cluster_floor(wrnn_engine* e, int32_t cluster_size, int32_t rounds, float* us);

/* Self-test of the tensor-core building blocks (TMA 128B-swizzle load, tcgen05.mma, TMEM load):
 * C[128][N] = A[128][512] * W[N][512]^T, fp16 bit patterns in, fp32 out, N in {16,32,48,64}.            */
int wrnn_debug_tc_gemm(wrnn_engine* e, const uint16_t* A, const uint16_t* W, int32_t N, float* C);
/* same on a CTA pair (tcgen05 cta_group::2, M = 256): A (256,512), W (N,512), N in {32,64,96,128}, C (256,N) */
int wrnn_debug_tc_gemm2(wrnn_engine* e, const uint16_t* A, const uint16_t* W, int32_t N, float* C);

/* number of kernel launches issued by this engine since creation (bench.py's gpu_launches)          */
int64_t wrnn_launch_count(const wrnn_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* WAVERNN_B200_H */
