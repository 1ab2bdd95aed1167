/*
 * iwq_b200.h -- C ABI of the B200 (sm_100a) weight fake-quant library `libiwq_b200.so`.
 *
 * The reference (LiuTielong/Iron_weight_only_quant) has no FFI layer on this path: the path sits
 * behind Python call signatures and is executed as sequences of ATen eager ops.  Each entry point
 * below therefore cites the reference *Python* interface whose arithmetic it replaces (file:line
 * relative to the reference repo); the Python modules in `iron_weight_only_quant_b200/` keep those
 * signatures and bind these symbols through ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain C types only; all data pointers are DEVICE pointers unless the name says `host`;
 *   - `stream` is a `cudaStream_t` passed as `void*`; work is enqueued on it, nothing synchronises
 *     unless stated; the caller makes `device` current or passes it explicitly where asked;
 *   - return value: 0 on success, a negative `IWQ_ERR_*` otherwise; `iwq_last_error()` gives text;
 *   - weights are row-major contiguous 2-D `[rows, cols]` of fp16 or bf16, `out` may alias `w`
 *     (the reference overwrites the weight storage in place: quant_linear.py:949).
 */
#ifndef IWQ_B200_H_
#define IWQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IWQ_ABI_VERSION 4

/* storage dtype of weights, scales and zeros */
enum { IWQ_F16 = 0, IWQ_BF16 = 1 };

/* quantisation scheme */
enum {
  IWQ_SCHEME_INT = 0,      /* quant_funcs.py:4-46 == quant_linear.py:885-956 (INT min-max, asym/sym)   */
  IWQ_SCHEME_FP_CODEC = 1, /* quant_linear.py:724-883 + _float_to_fp :126-163 + _fp_to_float :213-235 */
  IWQ_SCHEME_FP_RNE = 2,   /* fp4_quantize_cpu.py:37-72 (_fp_scale, ideal RNE minifloat, absmax scale) */
  IWQ_SCHEME_BFP = 3       /* quant_linear.py:648-723 (block floating point: shared exponent per group,
                              `bits` = w_bit: sign + min(w_bit-1, 11) mantissa bits; group > 0 only)   */
};

/* granularity: `group` > 0 -> contiguous runs of `group` elements along `cols`
 * (quant_funcs.py:10-12); IWQ_GROUP_PER_TENSOR (-1) -> one range for the matrix (quant_funcs.py:13-14,
 * quant_linear.py:898-900); IWQ_GROUP_PER_CHANNEL (-2) -> one range per row (quant_linear.py:901-903). */
enum { IWQ_GROUP_PER_TENSOR = -1, IWQ_GROUP_PER_CHANNEL = -2 };

enum {
  IWQ_OK = 0,
  IWQ_ERR_INVALID = -1,     /* bad argument (maps to ValueError)                                        */
  IWQ_ERR_INDIVISIBLE = -2, /* cols % group != 0 (maps to AssertionError, quant_funcs.py:11)            */
  IWQ_ERR_UNSUPPORTED = -3, /* combination the kernels do not implement (maps to NotImplementedError)   */
  IWQ_ERR_CUDA = -4,        /* a CUDA runtime call failed                                               */
  IWQ_ERR_ALIGN = -5,       /* pointer not 16-byte aligned                                              */
  IWQ_ERR_FP16_RANGE = -6   /* minifloat max not representable in fp16 (reference raises RuntimeError)  */
};

/* status bits OR-ed into the optional device status word by the kernels */
enum {
  IWQ_STATUS_NAN = 1,    /* some segment produced NaN: the condition of `assert torch.isnan(t).sum()==0`
                            (quant_funcs.py:40); checked lazily by the host */
  IWQ_STATUS_TIMEOUT = 2 /* iwq_fake_quant_per_tensor_fused: a peer's ranges did not arrive within the spin limit */
};

typedef struct iwq_config {
  int32_t dtype;     /* IWQ_F16 / IWQ_BF16                                                             */
  int32_t scheme;    /* IWQ_SCHEME_*                                                                   */
  int32_t bits;      /* INT: n_bits (1..15).  BFP: w_bit (2..15).  FP schemes: ignored                 */
  int32_t exp_bits;  /* FP schemes: exponent bits E (1..5)                                             */
  int32_t mant_bits; /* FP schemes: mantissa bits M (0..6)                                             */
  int32_t symmetric; /* 1: absmax scale, no zero point (zero_point=False); 0: min/max (zero_point=True)*/
  int32_t group;     /* > 0, IWQ_GROUP_PER_TENSOR, IWQ_GROUP_PER_CHANNEL                               */
  int32_t variant;   /* kernel variant: 0 = default (TMA bulk-copy staged), 1 = direct global loads    */
  int32_t quant_dim; /* 0: segments run along `cols` (default); 1: along `rows`, i.e. the reference's
                        quant_dim=1 "transpose, group, transpose back" (quant_linear.py:640-647); `group`
                        then divides `rows`, per-channel means one segment per COLUMN                   */
  int32_t approx;    /* FP_CODEC only: 0 exact decode; 1 aligned ("approximate") decode, _fp_decode_aligned
                        quant_linear.py:237-285; 2 double-approximate, :288-363 (quant_dim 1 only).  Both
                        force the symmetric absmax scale like quantize_weight_approximate (:470-632)     */
  int32_t hi_align_start;     /* approx: fp{4,6,8}_hi_align_start                                       */
  int32_t hi_align_exp_field; /* approx: fp{4,6,8}_hi_align_exp_field                                   */
  int32_t tail_pad_bits;      /* approx: fp{4,6,8}_tail_pad_bits (may be negative)                      */
  int32_t code_bits;          /* layout of iwq_matrix.codes: 0 or 8 = one byte per code; 4 = packed W4 format,
                                 two codes per byte, element 2i in the low nibble (INT bits <= 4, minifloats
                                 with 1 + E + M <= 4); no reference counterpart -- the reference discards its
                                 codes (quant_linear.py:613-625, :771), this is SURVEY section 8f row 4         */
} iwq_config_t;

typedef struct iwq_matrix {
  const void* w;  /* [rows, cols] input                                                               */
  void* out;      /* [rows, cols] fake-quantised output; may equal `w`                                */
  void* scales;   /* optional [G] scales in the weight dtype (G = rows*cols/group | rows | 1)         */
  void* zeros;    /* optional [G]: INT asym zero points / FP asym mid points, weight dtype            */
  void* codes;    /* optional [rows, cols] uint8 (or [rows, cols/2] when cfg.code_bits == 4): INT code
                     (two's complement for sym) / minifloat code; quant_dim = 0 only                 */
  int64_t rows;
  int64_t cols;
} iwq_matrix_t;

/* ---- library ---------------------------------------------------------------------------------- */
int iwq_abi_version(void);
const char* iwq_last_error(void);
/* bytes of device workspace `iwq_fake_quant_batched` needs for `n` matrices */
size_t iwq_workspace_bytes(int n);
/* Job-table cache.  A batched entry point copies its job table (64 bytes per matrix) into `workspace` on the caller's stream; that
 * copy has to wait for the previous kernel of the stream (~35 us of idle GPU between back-to-back launches).  enable = 1 declares
 * that the workspaces handed to this library are written by nobody else between calls: a table identical to the one last uploaded
 * to the same workspace on the same stream is then not uploaded again.  Off by default; the Python layer (which owns its
 * workspaces) turns it on.  No counterpart in the reference (quant_wrapper.py:52-82 launches per module). */
int iwq_workspace_cache(int enable);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* One fused pass over one matrix.  Replaces pseudo_quantize_tensor (quant_funcs.py:4-46), the INT
 * and FP4/6/8 branches of QuantLinear.quantize_weight (quant_linear.py:724-956) and
 * quantize_fp16_to_fp4_e1m2 (fp4_quantize_cpu.py:47-72).  `status` (device int32*, may be NULL)
 * receives IWQ_STATUS_* bits via atomicOr.  Per-tensor granularity runs a min/max pass first and needs
 * `range_ws` (device, 2 floats); other granularities ignore it. */
int iwq_fake_quant(const iwq_config_t* cfg, const iwq_matrix_t* m, float* range_ws, int32_t* status, void* stream);

/* The same for `n` matrices sharing one config, in ONE launch per pass (persistent work-list kernel):
 * replaces the per-module loop of quantize_model (quant_wrapper.py:52-82).  `workspace` is device
 * memory of at least iwq_workspace_bytes(n). */
int iwq_fake_quant_batched(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, void* workspace, size_t workspace_bytes,
                           int32_t* status, void* stream);

/* Per-tensor granularity (quant_funcs.py:13-14, quant_linear.py:898-900) for `n` matrices that are WHOLE on this GPU:
 * the library orders range passes and apply passes itself (per L2-sized slab of matrices, so that the apply pass
 * reads its slab from L2).  `range`: device float[2*n] scratch, receives {max|absmax, -min} per matrix. */
int iwq_fake_quant_per_tensor_batched(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, float* range, void* workspace,
                                      size_t workspace_bytes, int32_t* status, void* stream);

/* ---- per-tensor granularity on a matrix that is sharded across GPUs (no reference counterpart;
 *      SURVEY.md section 8e): range pass, (NCCL MAX all-reduce of `range` by the caller), apply pass.
 * `range`: device float[2*n] laid out {max_or_absmax, -min} per matrix, so one MAX reduction serves both. */
int iwq_minmax(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, float* range, void* workspace, size_t workspace_bytes,
               void* stream);
int iwq_fake_quant_with_range(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, const float* range, void* workspace,
                              size_t workspace_bytes, int32_t* status, void* stream);

/* The same three steps in ONE launch (pertensor_fused_kernel): range pass, in-kernel exchange of the ranges
 * through peer-mapped memory (NVLink / NVSwitch P2P stores + system-scope flags; no NCCL call) and apply pass,
 * software-pipelined over L2-sized batches of matrices so that the apply pass reads the weights from L2
 * (HBM traffic 4 B/weight instead of 6).  `peers` == NULL: single GPU (world 1), scratch comes from `workspace`.
 * Otherwise every rank of the group calls this with its own shards and the same `n`, shapes and config:
 *   slot_ptrs[r] / flag_ptrs[r] = rank r's exchange buffers AS MAPPED ON THIS GPU (e.g. torch symmetric memory),
 *   slots: float[2][world][2 * n] (double-buffered by epoch parity), flags: uint32[world][n] (>= one per batch),
 *   zero-initialised once; `epoch` must be non-zero and different from the previous call's on these buffers.
 * `workspace`: device memory of at least iwq_fused_workspace_bytes(n).  Status: IWQ_STATUS_NAN / IWQ_STATUS_TIMEOUT. */
typedef struct iwq_peers {
  int32_t world, rank;
  void* const* slot_ptrs; /* host array [world] of device pointers */
  void* const* flag_ptrs; /* host array [world] of device pointers */
  uint32_t epoch;
  uint32_t spin_limit;    /* 0 = default (about one second) */
} iwq_peers_t;
size_t iwq_fused_workspace_bytes(int n);
long long iwq_debug_fused_trace(int* n_batches); /* diagnostics: see iwq_host.cu */
/* dispatch decision for a [rows, cols] matrix under `cfg` (kernel, tile shape; no CUDA call): out = int32[8], see iwq_host.cu */
int iwq_debug_plan(const iwq_config_t* cfg, int64_t rows, int64_t cols, int32_t* out);
/* diagnostics: set a tuning knob (the IWQ_* environment variables, name without the prefix in lower case) at run time */
int iwq_debug_set_knob(const char* name, int value);
int iwq_fake_quant_per_tensor_fused(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, void* workspace, size_t workspace_bytes,
                                    const iwq_peers_t* peers, int32_t* status, void* stream);

/* ---- element-wise pieces, exposed for exhaustive parity tests ---------------------------------- */
/* decode(encode(x)) of the quant_linear minifloat codec on `n` 16-bit values (quant_linear.py:126-163,
 * :213-235); `codes` optional. flavour = IWQ_SCHEME_FP_CODEC or IWQ_SCHEME_FP_RNE (then: _fp_scale with S=1). */
int iwq_fp_roundtrip(int dtype, int flavour, int exp_bits, int mant_bits, const void* x, void* y, void* codes, int64_t n, void* stream);
/* _fp_scale (fp4_quantize_cpu.py:37-44) on `n` fp16 values with an ELEMENT-WISE fp16 scale array (the caller broadcasts
 * the reference's S): y = RNE-onto-the-(E, M)-grid(clamp(x / S, +-max_float)), not multiplied back by S. */
int iwq_fp_scale(int exp_bits, int mant_bits, const void* x, const void* scale, void* y, int64_t n, void* stream);
/* exhaustive device check of the reciprocal+FMA quotient against IEEE division for every (w, s) pair of
 * the dtype; writes the number of mismatching pairs to *mismatches (device int64*, caller zeroes it). */
int iwq_selftest_division(int dtype, unsigned long long* mismatches, void* stream);

/* ---- host-buffer pipeline (end-to-end path) ---------------------------------------------------- */
/* Fake-quantise a matrix that lives in HOST memory (pinned for full speed): chunks are copied to the
 * device, quantised and copied back on alternating streams so that H2D, kernel and D2H overlap.
 * `host_out` may equal `host_w`.  Synchronises before returning.  `device` is the CUDA device ordinal.
 * Only group > 0 and per-channel granularity (chunk boundaries fall on rows). */
int iwq_fake_quant_host(const iwq_config_t* cfg, const void* host_w, void* host_out, int64_t rows, int64_t cols, int device,
                        size_t chunk_bytes, int32_t* host_status);

/* The same for `n` host matrices in ONE pipeline: chunks of consecutive matrices follow each other through the
 * staging ring, which is drained once at the end instead of once per matrix (a [1024, 8192] matrix alone is a
 * single chunk, i.e. no H2D/D2H overlap at all).  This is what a caller holding a whole state dict in pinned
 * memory uses; replaces the per-module loop of quantize_model (quant_wrapper.py:52-82) for host-resident weights. */
typedef struct iwq_host_matrix {
  const void* host_w; /* [rows, cols] 16-bit weights in host memory (pinned for full speed) */
  void* host_out;     /* may equal host_w                                                    */
  int64_t rows;
  int64_t cols;
  void* host_scales;  /* optional [G] scales in the weight dtype, HOST memory (G = rows*cols/group | rows): what the module
                         form stores in `self.scales` (quant_linear.py:924-932, :761-768); NULL = not wanted        */
  void* host_zeros;   /* optional [G] zero points / mid points, HOST memory; NULL = not wanted                          */
} iwq_host_matrix_t;
int iwq_fake_quant_host_batched(const iwq_config_t* cfg, const iwq_host_matrix_t* ms, int n, int device, size_t chunk_bytes,
                                int32_t* host_status);

#ifdef __cplusplus
}
#endif
#endif /* IWQ_B200_H_ */
