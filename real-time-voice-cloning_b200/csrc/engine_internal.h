// engine_internal.h -- declarations shared by the engine's translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace wrnn {

constexpr int kUnitsF32 = 4;              // hidden units per CTA in the fp32 loop
constexpr int kCtasF32 = kRnn / kUnitsF32;  // 128 CTAs, one per SM (148 available)
constexpr int kMaxFoldsPerLaunch = 256;

// Arguments of the persistent loop kernels.
struct LoopParams {
    // GEMV weights, row-major [rows][512] fp32 (device)
    const float* Whh1;    // rnn1.weight_hh_l0            [1536][512]
    const float* Wih2a;   // rnn2.weight_ih_l0[:, :512]   [1536][512]
    const float* Whh2;    // rnn2.weight_hh_l0            [1536][512]
    const float* Wfc1a;   // fc1.weight[:, :512]          [512][512]
    const float* Wfc2a;   // fc2.weight[:, :512]          [512][512]
    const float* Wfc3;    // fc3.weight                   [C][512]
    const float* v1;      // rnn1.weight_ih_l0 @ I.weight[:, 0]          [1536]  (rank-1 term of the previous sample)
    const float* v2;      // rnn2.weight_ih_l0[:, :512] @ I.weight[:, 0] [1536]
    const float* v3;      // fc1.weight[:, :512] @ I.weight[:, 0]        [512]
    const float* bhn1;    // rnn1.bias_hh_l0[1024:]       [512]
    const float* bhn2;    // rnn2.bias_hh_l0[1024:]       [512]
    const float* bfc3;    // fc3.bias                     [C]
    // per-frame conditioning tables written by the front end (cond.cu)
    const float4* TA1;    // [rows][512] {gi1_r, gi1_z, gi1_n, fc1}   aux part + biases (incl. the I layer's share)
    const float4* TA2;    // [rows][512] {gi2_r, gi2_z, gi2_n, fc2}
    const float4* TQ1;    // [rows][512] {gi1_r, gi1_z, gi1_n, fc1}   share of one padded mel frame
    const float4* TQ2;    // [rows][512] {gi2_r, gi2_z, gi2_n, 0}
    const float* coef;    // [200][kTaps] upsample interpolation weights (phase, padded-frame tap)
    const FoldDesc* folds;
    int B, S, C, Cpad, CR, FB, mode;
    unsigned long long seed;
    // exchange buffers, {value, tag} words
    unsigned long long *bH1, *bH2, *bF1, *bF2;               // [B][512]
    unsigned long long* bLG;                                  // [B][Cpad]
    unsigned long long* bX;                                   // [B]
    float* samples;          // [B][S] value fed back at each step
    float* logits_out;       // optional [B][S][C]
    const float* forced;     // optional [B][S]
    int* progress;           // mapped host int, written every 100 steps
    int* abort_flag;         // device int
};

size_t loop_f32_smem_bytes(int B, int FB, int CR);
int loop_f32_pick_fb(int B, int CR, size_t smem_limit);
cudaError_t launch_loop_f32(const LoopParams& p, cudaStream_t stream);
cudaError_t set_spin_deadline(long long cycles);

// ---- tensor-core loop (loop_tc.cu) -------------------------------------------------------------------------
constexpr int kTcGroups = 2;      // independent groups of CTAs, each with a full fp16 copy of the loop weights
constexpr int kTcCtas = 64;       // CTAs per group
constexpr int kTcUnits = 8;       // hidden units per CTA
constexpr int kTcSets = 4;        // fold sets (<= 128 folds each) a group pipelines through its CTAs
constexpr int kTcMaxFolds = kTcGroups * kTcSets * 128;
#ifndef WRNN_EXPAND_STEPS
#define WRNN_EXPAND_STEPS 32
#endif
constexpr int kExpandSteps = WRNN_EXPAND_STEPS;   // steps per block of expand_cond = granularity at which the loop may follow the expansion
constexpr int kTcKbPerOp = 2;      // k-blocks (64 columns each) one TMA operation of the loop brings in

struct TcParams {
    const unsigned char* wimg;   // [kTcCtas][loop_tc_weight_image_bytes()] per-CTA weight images in shared-memory layout
    const unsigned char* wimg_s; // RAW sampler CTAs: [4][loop_tc_raw_sampler_image_bytes()] quarters of fc3 as N=128 tiles
    int raw_samplers;            // != 0: RAW, 512 classes: fc3 + the draw run on 4 dedicated sampler CTAs per group
    const float *v1, *v2, *v3, *bhn1, *bhn2, *bfc3;
    const float4* CS;            // per-sample conditioning [virtual group][step][row < Mg][256 unit pairs][4 float4] (expand_cond)
    int Mg;                      // folds per virtual group (= group x set); fold f is row f % Mg of virtual group f / Mg
    int nsets;                   // fold sets per group in this launch (1..kTcSets)
    int pair;                    // != 0: unit-owning CTAs work as CTA pairs (cta_group::2); nsets is 2 or 4
    const FoldDesc* folds;
    int B, S, C, Cpad, mode;
    // conditioning expansion inside the loop kernel (expander CTAs on the SMs the loop leaves free): per-frame tables in,
    // CS out, one completion counter per 16-step chunk; cs_done == nullptr: CS was expanded before the launch
    unsigned int* cs_done;
    unsigned int* cs_consumed;   // per chunk: unit-owning CTAs that have read it (the expanders reuse CS as a ring of cs_steps steps)
    int cs_steps;                // steps CS holds per fold (== S rounded up to a chunk when it is not a ring)
    float4* CSw;
    const float4 *TA1, *TA2, *TQ1, *TQ2;
    const float* coef;
    int n_expanders;
    int tile_bytes;              // bytes one k-block of a TMA operation occupies: box_rows * 128
    int flags;                   // bit 0: epilogue warps release the counters themselves (no publisher warp); 1: 4-slot ring; 2: L2 prefetch of the next record
    unsigned long long seed;
    __half *H1, *H2, *F1, *F2;   // activation exchange, [kTcGroups*kTcSets*128][512] fp16
    unsigned int* counters;      // [kTcGroups*kTcSets][4] arrival counters (monotonic)
    unsigned long long* bLG;     // [kTcGroups*kTcSets*128][Cpad] logits exchange words (RAW); RAW samplers: [2][rows][4] {max, sum} pairs
    unsigned long long* bX;      // [kTcGroups*kTcSets*128] sample exchange words
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
    int* abort_flag;
    long long* trace;            // optional [16 steps][32 slots] SM-clock timeline of CTA 0 (WRNN_TC_TRACE=1)
};
size_t loop_tc_weight_image_bytes();
size_t loop_tc_raw_sampler_image_bytes();
int loop_tc_raw_sampler_ctas();
int loop_tc_sampler_ctas(int mode, int raw_samplers, int pair);
bool loop_tc_pair_fits(int nsets, int grid);             // the pair kernel's 2-CTA clusters can all be co-resident on this device   // CTAs past the unit-owning groups that run fc3 + the draw
cudaError_t set_tc_deadline(long long cycles);
cudaError_t launch_loop_tc(const TcParams& p, const void* tmaps, cudaStream_t stream);
cudaError_t launch_expand_cond(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                               const FoldDesc* folds, int B, int S, int Mg, float4* CS, cudaStream_t stream);

// ---- role-specialised tensor-core loop, MOL, <= 128 folds per group (loop_rs.cu) ------------------------------------
#ifndef WRNN_RS_FC_CTAS
#define WRNN_RS_FC_CTAS 8
#endif
constexpr int kRsT1 = 16, kRsT2 = 16, kRsT3 = WRNN_RS_FC_CTAS, kRsT4 = WRNN_RS_FC_CTAS;     // CTAs per role: GRU1 (+fc3 + draw), GRU2, fc1, fc2
constexpr int kRsCtas = kRsT1 + kRsT2 + kRsT3 + kRsT4;          // one group (MOL); RAW adds n_samplers sampler CTAs (T5) per group
constexpr int kRsMaxSamplers = 8;                               // RAW: sampler CTAs per group; each holds C / 8 classes of fc3 (64 or 128)
constexpr int kRsBufs = 2;                                      // exchange matrices are double-buffered by step parity
constexpr int kRsMaxFoldsPerGroup = 128;
constexpr int kRsChunk = 4;                                     // steps per chunk of conditioning records (granularity of the ring's counters)
struct RsParams {
    const unsigned char *w1, *w2, *w3, *w4;   // per-role weight images [CTA][loop_rs_image_bytes(role)], shared-memory layout
    const unsigned char* w5;     // RAW: sampler images [n_samplers][128 classes of fc3 as an N = 128 tile]
    int mode, C, n_samplers;     // WRNN_MODE_*; classes; RAW: sampler CTAs per group (8), MOL: 0
    int qcols;                   // RAW: classes per sampler CTA = C / n_samplers (64 or 128)
    int ctas;                    // CTAs per group = kRsCtas + n_samplers
    unsigned long long* bP;      // RAW: soft-max partials {max, sum} as tagged words [G][2 (step parity)][128 folds][n_samplers][2]
    const float *v1, *v2, *v3, *bhn1, *bhn2, *bfc3;
    const float* CS;             // per-sample conditioning [group][cs_steps][Ng][8][512] fp32 (expand_cond_rs_kernel)
    int cs_steps;                // steps the ring holds (a multiple of kRsChunk; == the padded S when everything is expanded up front)
    // expander CTAs past the groups (cs_done != nullptr): per-frame tables in, records out, one produced / consumed counter per chunk
    unsigned int *cs_done, *cs_consumed;
    float* CSw;
    const float4 *TA1, *TA2, *TQ1, *TQ2;
    const float* coef;
    int n_expanders;
    // inline conditioning (inl != 0; no records, no expanders): the aux share + biases of a record as one per-FRAME row
    // FR[frame row][8][512] fp32, the mel share as one more K = 80 slab of the role's on-path product -- operand M16[row][80] fp16
    // (the upsampled mel of every sample, row = (tq_row0 + frame) * 200 + phase; row m16_zero is all zeros), weights wx1..3
    // (per-CTA tiles [k-block 2][rows][128 B] of W_ih1 / W_ih2a / fc1a . I[:, mel]: 128 / 96 / kFU rows)
    int inl;
    const unsigned char *wx1, *wx2, *wx3;
    const float* FR;
    const __half* M16;
    long long m16_zero;
    unsigned int* place;         // optional [grid] zeroed words: logical CTA index from the physical SM (rank of %smid), rotated by rot
    int rot;
    int canary_all;              // phase 1 of an ingest waits for all producers' canaries (1) or the first one (0)
    int offpath_delay_ns;        // T1 waits this long before it reads h1(t) for the recurrent product (the T2 CTAs read it first)
    int Ng, G;                   // folds per group (<= 128), groups; fold f = row f % Ng of group f / Ng
    const FoldDesc* folds;
    int B, S;
    unsigned long long seed;
    uint4* X;                    // exchange: [G][5 matrices][kRsBufs][64 chunks][128 folds] 16-byte chunks; all bytes 0xFF at launch
    unsigned long long* bX;      // [G][128] sample words {value, tag}; zero at launch
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
    int* abort_flag;
    int* dbg;                    // optional mapped host memory [CTA][32] checkpoints (WRNN_RS_DEBUG=1)
    unsigned long long* trace;   // optional [CTA][8 steps][16 events] %globaltimer stamps (WRNN_RS_TRACE=path)
};
size_t loop_rs_image_bytes(int role);
size_t loop_rs_exchange_bytes(int groups);
cudaError_t set_rs_deadline(long long cycles);
cudaError_t launch_loop_rs(const RsParams& p, cudaStream_t stream);
struct UttDesc;
size_t loop_rs_ximage_bytes(int role);
cudaError_t launch_rs_inline_tables(const float4* TA1, const float4* TA2, const float* mel, const UttDesc* utts, int n_utts, const float* coef,
                                    int rows, float* FR, __half* M16, cudaStream_t stream);
cudaError_t launch_expand_cond_rs(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                                  const FoldDesc* folds, int B, int S, int Ng, int cs_steps, float* CS, cudaStream_t stream);

// ---- runtimeracer-wavernn topology (loop_rr.cu): four GRU-256 + five FC layers, fp32 ---------------------------------
constexpr int kRrH = 256;            // rnn_dims = fc_dims (config/hparams.py:363-364)
constexpr int kRrCtas = 128;         // 2 hidden units of every layer per CTA
constexpr int kRrMaxFolds = 64;      // folds per launch (longer batches run in waves): activations + running sums = 2 KB of shared memory per fold
struct RrLoopParams {
    const float* Whh[4];         // rnn1..4 weight_hh [768][256]
    const float* Wih[3];         // rnn2, rnn3[:, :256], rnn4 weight_ih [768][256] (rnn1's is folded into the tables)
    const float *M12, *M34;      // fc2 fc1[:, :256], fc4 fc3[:, :256]  [256][256] (no activation between the pairs)
    const float* Wfc5;           // [C][256]
    const float* u;              // [13][256] coefficients of the previous sample: GRU1..4 (r, z, n) and the M12 station
    const float* bhn;            // [4][256] b_hn of the four GRUs
    const float* bfc5;           // [C]
    // per-frame tables, 4 float4 per hidden unit: {c1 r,z,n, c5} {c2 r,z,n, c6} {c3 r,z,n, 0} {c4 r,z,n, 0}
    // TA[frame row][256][4] from the aux channels (+ biases), TQ[padded frame row][256][4] from the padded mel frames
    const float4 *TA, *TQ;
    const float* coef;           // [200][kTaps] interpolation weights of the upsampling stack
    const FoldDesc* folds;
    int B, S, C, Cpad, CR, mode;
    unsigned long long seed;
    unsigned long long* bH[4];   // exchange words {value, tag}: [B][256] each
    unsigned long long *bY2, *bY4, *bLG, *bX;
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
    int* abort_flag;
};
using RrParams = RrLoopParams;
size_t loop_rr_smem_bytes(int B, int CR);
cudaError_t set_rr_deadline(long long cycles);
cudaError_t launch_loop_rr(const RrLoopParams& p, cudaStream_t stream);

// ---- geneing-wavernn topology (loop_gn.cu): one GRU-256, fc1 (-> 128, ReLU), fc3 (-> classes), fp32 -------------------------
constexpr int kGnH = 256;            // rnn_dims (config/hparams.py:295)
constexpr int kGnFc = 128;           // fc_dims (:296)
constexpr int kGnAux = 32;           // res_out_dims / 2 (:298, geneing_version.py:103)
constexpr int kGnCh = 64;            // compute_dims = res_out_dims (:297-298)
constexpr int kGnResBlocks = 3;      // :299
constexpr int kGnCtas = 128;         // 2 GRU units + 1 fc1 unit per CTA
constexpr int kGnMaxFolds = 96;      // folds per launch
struct GnLoopParams {
    const float* Whh;            // rnn1 weight_hh [768][256]
    const float* Wfc1a;          // fc1[:, :256] [128][256]
    const float* Wfc3;           // [C][128]
    const float *u1, *u2;        // coefficients of the previous sample: GRU1 gates [3][256], fc1 [128]
    const float *bhn, *bfc3;     // [256], [C]
    // per-frame tables, one float4 per GRU unit j: {c1 r, z, n of unit j, c2 of fc1 unit j/2 (even j; 0 for odd j)}
    const float4 *TA, *TQ;       // [frame row][256], [padded frame row][256]
    const float* coef;
    const FoldDesc* folds;
    int B, S, C, Cpad, CR, mode;
    unsigned long long seed;
    unsigned long long *bH, *bF, *bLG, *bX;   // exchange words: [B][256], [B][128], [B][Cpad], [B]
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
    int* abort_flag;
};
size_t loop_gn_smem_bytes(int B, int CR);
cudaError_t set_gn_deadline(long long cycles);
cudaError_t launch_loop_gn(const GnLoopParams& p, cudaStream_t stream);

// ---- cluster-local tensor-core loop, MOL (loop_tc2.cu) --------------------------------------------------------------
struct Tc2Params {
    const unsigned char* wimg;   // [16][loop_tc2_image_bytes()] per-CTA streams of pre-swizzled weight tiles
    int img_bytes;
    const float *v1, *v2, *v3, *bhn1, *bhn2, *bfc3;
    const float* CS;             // per-sample conditioning [cluster][step][16 CTAs][4 gates][32 units][32 folds][2] (expand_cond2)
    const FoldDesc* folds;
    int B, Bc, S;
    unsigned long long seed;
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
    int* abort_flag;
    long long* trace;
};
size_t loop_tc2_image_bytes();
cudaError_t set_tc2_deadline(long long cycles);
int loop_tc2_max_clusters();
cudaError_t launch_loop_tc2(const Tc2Params& p, int n_clusters, cudaStream_t stream);
cudaError_t launch_expand_cond2(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                                const FoldDesc* folds, int B, int Bc, int n_clusters, int S, float* CS, cudaStream_t stream);

// ---- block-sparse cluster loop (loop_sparse.cu) -----------------------------------------------------------------
struct SparseParams {
    const unsigned char* wimg;   // [cluster size][img_stride] per-CTA images: int header[16] (byte offsets of rowptr/col/w per
    int img_stride;              // stage at header[4+3s..]), then rowptr (int), group columns (u8), weights (float4)
    const float *v1, *v2, *v3, *bhn1, *bhn2, *bfc3;
    const float4 *TA1, *TA2, *TQ1, *TQ2;
    const float* coef;
    const FoldDesc* folds;
    int B, Bc, S, C, Cpad, CRs, mode;
    unsigned long long seed;
    float* samples;
    float* logits_out;
    const float* forced;
    int* progress;
};
size_t loop_sparse_smem_bytes(int cluster, int img_stride, int Bc, int Cpad, int CRs);
cudaError_t launch_loop_sparse(const SparseParams& p, int cluster, int n_clusters, cudaStream_t stream);

// ---- conditioning front end (cond.cu) ---------------------------------------------------------------
struct UttDesc {
    long long mel_off;   // float offset of this utterance's (80,T) block in the device mel buffer
    int T;
    int ta_row0;         // first of T+1 rows in the aux-row space
    int tq_row0;         // first of T+4 rows in the padded-frame space
    int pad_;
};
// X0[ta_row][400] = im2col of the zero-padded mel for conv_in (k=5); MP[tq_row][80] = padded mel, time-major
cudaError_t launch_im2col(const float* mel, const UttDesc* utts, int n_utts, int ta_rows, int tq_rows,
                          float* X0, float* MP, cudaStream_t stream);
// C[M][N] = act(A[M][K] W[N][K]^T + bias) (+ R)   ; fp32 ; K % 16 == 0 ; N % 64 == 0
cudaError_t launch_gemm_f32(const float* A, const float* W, const float* bias, const float* R, float* C,
                            int M, int N, int K, int relu, cudaStream_t stream);
cudaError_t launch_zero_rows(float* aux, const UttDesc* utts, int n_utts, cudaStream_t stream, int channels = 128);

// ---- conditioning contractions on tensor cores (cond_tc.cu) ------------------------------------------------
struct GemmTcArgs {
    int M, N;            // output rows (frames) and columns (multiple of 128)
    int nkb;             // K blocks of 64 columns
    int kb_per_tap;      // K blocks per convolution tap (== nkb for a plain GEMM)
    int row_shift;       // A rows advance by row_shift per tap (k=5 convolution without im2col)
    int relu;
    float scale;         // applied to the accumulator before the bias (weights are pre-scaled by 1/scale)
    const float* bias;   // [N] or null
    const float* R;      // residual [M][N] fp32 or null
    const float* rowmask;// [M] or null: output rows are multiplied by it (bias-only rows of an utterance -> 0)
    float* C;            // fp32 output [M][N] or null
    __half* Chi;         // hi/lo fp16 output [M][N] or null (operand of the next layer)
    __half* Clo;
    int* status;
};
cudaError_t launch_gemm_tc_split(const __half* Ahi, const __half* Alo, int Arows, int Acols, const __half* Whi, const __half* Wlo,
                                 const GemmTcArgs& args, cudaStream_t stream);
cudaError_t launch_mel_split(const float* mel, const UttDesc* utts, int n_utts, int rows, __half* hi, __half* lo, float* rowmask,
                             cudaStream_t stream);

// ---- post chain (post.cu) ----------------------------------------------------------------------------
struct PostUtt {
    long long samp_off;   // float offset of this utterance's (F,S) block in the samples buffer
    long long wav_off;    // double offset of its output in the wav buffer
    int F;                // folds (1 when not batched)
    int wave_len;         // (T-1)*200
};
cudaError_t launch_post(const float* samples, const PostUtt* utts, int n_utts, int max_wave_len, int S, int batched,
                        int target, int overlap, const double* fade_in, const double* fade_out, int mu_law, int n_classes,
                        int preemph, double* scratch, double* wav, cudaStream_t stream);

cudaError_t launch_xfade_unfold_f64(const double* y, int F, int S, int overlap, const double* fade_in,
                                    const double* fade_out, long long total_len, double* out, cudaStream_t stream);

// ---- tensor-core building blocks (tma_host.cu, tc_gemm_test.cu) ----------------------------------------
cudaError_t make_tmap_f16_2d(void* tmap_out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                             uint32_t box_cols);
cudaError_t make_tmap_f16_kblocks(void* tmap_out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_kb);
cudaError_t run_umma_rate(int N, int iters, int mode, long long* out_dev, cudaStream_t stream);
cudaError_t run_tc_gemm2_test(const void* A_dev, const void* W_dev, int N, float* C_dev, int* status_dev, cudaStream_t stream);
cudaError_t run_tc_gemm_test(const void* A_dev, const void* W_dev, int N, float* C_dev, int* status_dev, cudaStream_t stream);

// ---- exchange-floor microbenchmark (bench_floor.cu) -----------------------------------------------------
cudaError_t launch_floor_ll(unsigned long long* buf, int rounds, int* abort_flag, cudaStream_t stream);
cudaError_t launch_floor_cluster(int cluster_size, int rounds, float* sink, cudaStream_t stream);
cudaError_t launch_floor_counter(unsigned int* counter, float* data, int rounds, int* abort_flag, cudaStream_t stream);

}  // namespace wrnn
