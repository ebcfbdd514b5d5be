/* gimb200.h - C ABI of libgimb200.so: the B200 (sm_100a) gim_loftr dense-matching hot path.
 *
 * The reference (xuelunshen/gim) is pure Python/PyTorch and has no FFI layer; its boundary for this
 * path is the nn.Module API.  Each entry point below therefore cites the reference *Python* interface
 * it stands in for.  The Python shim `gim_b200/loftr.py::LoFTR` keeps the module signature and calls
 * these functions through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success, non-zero on error; the message is in gimb_last_error()
 *     (thread-local); nothing throws across the ABI;
 *   - device pointers are CUDA device memory on the handle's device; the caller owns every buffer
 *     (inputs, outputs, workspace); the library owns only its packed weights;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - gimb_loftr_forward() performs ONE host synchronisation: the read-back of the match count M
 *     that sizes the fine stage - the same point where the reference synchronises in
 *     `torch.where(mask_v)` (networks/loftr/utils/coarse_matching.py:193).
 */
#ifndef GIMB200_H
#define GIMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIMB_ABI_VERSION 1

/* ---- packed weight blob (built by gim_b200/weights.py::pack_loftr_blob) ------------------------
 * header | n_entries * entry | pad to 256 | tensor data (each tensor 256-byte aligned, fp32).   */
#define GIMB_BLOB_MAGIC 0x31304257424D4947ULL /* "GIMBWB01" little endian */
typedef struct {
  uint64_t magic;
  uint32_t version;
  uint32_t n_entries;
  uint64_t data_offset; /* from blob start */
  uint64_t total_bytes;
} gimb_blob_header;

typedef struct {
  char name[96];
  uint32_t ndim;
  uint32_t shape[4];
  uint32_t reserved;
  uint64_t offset; /* from data_offset */
  uint64_t nbytes;
} gimb_blob_entry;

/* ---- matcher configuration: the subset of `config['match_coarse']` / `config['fine_window_size']`
 * that LoFTR.__init__ consumes (networks/loftr/loftr.py:15-29, networks/loftr/config.py:7-46). */
typedef struct {
  float thr;               /* match_coarse.thr               (0.2) */
  int32_t border_rm;       /* match_coarse.border_rm         (2)   */
  float dsmax_temperature; /* match_coarse.dsmax_temperature (0.1) */
  int32_t fine_window;     /* fine_window_size               (5)   */
} gimb_loftr_cfg;

/* ---- outputs of one forward: the keys LoFTR.forward adds to `data`
 * (networks/loftr/utils/coarse_matching.py:229-257, networks/loftr/utils/fine_matching.py:59-72).
 * All arrays are DEVICE buffers with room for `capacity` rows; rows [0, M) are valid and ordered
 * by (b, i) ascending exactly like torch.where (coarse_matching.py:192-195).                   */
typedef struct {
  int64_t capacity;  /* in: rows available; must be >= N * min(L, S)                            */
  int64_t* b_ids;    /* [M]   int64  == m_bids                                                  */
  int64_t* i_ids;    /* [M]   int64  coarse cell index in image0 (row-major h0c x w0c)          */
  int64_t* j_ids;    /* [M]   int64  coarse cell index in image1                                */
  float* mconf;      /* [M]   fp32   conf_matrix[b, i, j]                                       */
  float* mkpts0_c;   /* [M,2] fp32   (x, y) pixels                                              */
  float* mkpts1_c;   /* [M,2] fp32                                                              */
  float* mkpts0_f;   /* [M,2] fp32   == mkpts0_c                                                */
  float* mkpts1_f;   /* [M,2] fp32   sub-pixel refined                                          */
  float* expec_f;    /* [M,3] fp32   (ex, ey, std)                                              */
} gimb_loftr_out;

/* ---- optional debug taps: when a pointer is non-NULL the intermediate tensor is copied there.
 * Layouts are the library's own (channels-last).  Used by the stage-level parity tests.          */
typedef struct {
  float* feat_c_backbone0; /* [N, H0/8, W0/8, 256]  FPN coarse map of image0 before PE           */
  float* feat_c_backbone1; /* [N, H1/8, W1/8, 256]                                               */
  float* feat_f0;          /* [N, H0/2, W0/2, 128]  FPN fine map of image0                       */
  float* feat_f1;          /* [N, H1/2, W1/2, 128]                                               */
  float* feat_c0;          /* [N, L, 256]  after the coarse transformer                          */
  float* feat_c1;          /* [N, S, 256]                                                        */
  float* fine_win0;        /* [capacity, 25, 128]  after the fine transformer                    */
  float* fine_win1;        /* [capacity, 25, 128]                                                */
  float* conf_matrix;      /* [N, L, S]  dual-softmax confidence (`data['conf_matrix']`)         */
} gimb_loftr_taps;

typedef struct gimb_loftr gimb_loftr;

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* gimb_last_error(void);
int gimb_abi_version(void);

/* Replaces: LoFTR.__init__ + load_state_dict + .eval().to(device)
 * (networks/loftr/loftr.py:15-41, 93-99; demo.py:335,373-400).  `blob` is HOST memory. */
int gimb_loftr_create(const void* blob, size_t nbytes, const gimb_loftr_cfg* cfg, int device,
                      gimb_loftr** out);
void gimb_loftr_destroy(gimb_loftr* h);

/* Bytes of device workspace gimb_loftr_forward needs for a batch of n pairs of the given sizes
 * (h, w multiples of 8).  No reference counterpart (PyTorch allocates implicitly). */
int gimb_loftr_workspace_bytes(gimb_loftr* h, int n, int h0, int w0, int h1, int w1, size_t* bytes);

/* Replaces: the `pe` buffer PositionEncodingSine registers at construction
 * (networks/loftr/utils/position_encoding.py:22-37, used at loftr.py:74-75).  Uploads and caches the
 * table for coarse maps of hc x wc cells: host_pe is [hc*wc, 256] fp32, token-major
 * (pe[(y*wc + x), c] = reference pe[0, c, y, x]).  Must be called once per coarse size before forward. */
int gimb_loftr_set_pe(gimb_loftr* h, int hc, int wc, const float* host_pe);

/* Replaces: LoFTR.forward(data) (networks/loftr/loftr.py:43-91) for DEVICE inputs.
 *   color0 [n,3,h0,w0], color1 [n,3,h1,w1]  fp32 NCHW RGB in [0,1]   (data['color0'|'color1'])
 *   mask0 [n,h0/8,w0/8], mask1 [n,h1/8,w1/8] uint8 0/1, both NULL or both set (data['mask0'|'mask1'])
 *   scale0, scale1 [n,2] fp32 (w,h) factors, both NULL or both set      (data['scale0'|'scale1'])
 *   m_out (host): number of matches M.  */
int gimb_loftr_forward(gimb_loftr* h, const float* color0, const float* color1,
                       const uint8_t* mask0, const uint8_t* mask1, const float* scale0,
                       const float* scale1, int n, int h0, int w0, int h1, int w1, void* workspace,
                       size_t workspace_bytes, const gimb_loftr_out* out,
                       const gimb_loftr_taps* taps, int64_t* m_out, void* stream);

/* Same call for HOST buffers (pinned or pageable): inputs are copied host->device into `dev_inputs`,
 * the forward runs, and rows [0, M) of every non-NULL array of `host_out` are copied back
 * (host_out->capacity rows available).  `dev_out` supplies the device result arrays.  This is the
 * end-to-end entry the ZEB harness / demo.py path maps to (trainer/lightning.py:158-159 moves the
 * batch, runs the model; tools/metrics.py:125-130 reads the results back).
 * *h2d_bytes / *d2h_bytes report the traffic of this call. */
int gimb_loftr_forward_host(gimb_loftr* h, const float* color0, const float* color1,
                            const uint8_t* mask0, const uint8_t* mask1, const float* scale0,
                            const float* scale1, int n, int h0, int w0, int h1, int w1,
                            void* dev_inputs, size_t dev_inputs_bytes, void* workspace,
                            size_t workspace_bytes, const gimb_loftr_out* dev_out,
                            const gimb_loftr_out* host_out, int64_t* m_out, uint64_t* h2d_bytes,
                            uint64_t* d2h_bytes, void* stream);
/* Bytes of device staging gimb_loftr_forward_host needs for the inputs (`dev_inputs`). */
int gimb_loftr_host_staging_bytes(int n, int h0, int w0, int h1, int w1, int with_mask,
                                  int with_scale, size_t* bytes);

/* GPU pre-processing entry.  Replaces, per image, the host work of the ZEB loader after cv2.resize
 * (datasets/utils.py:112-124: zero padding at the bottom / right, float / 255, HWC -> CHW, padding mask) and of
 * demo.py:166-176: img0 [n, ih0, iw0, 3] / img1 [n, ih1, iw1, 3] are uint8 RGB HWC HOST buffers; they are copied as
 * bytes (a quarter of the fp32 traffic), converted and zero-padded to (h0, w0) / (h1, w1) on the device, and when any
 * padding happens mask0 / mask1 are generated at 1/8 resolution exactly like datasets/kitti/kitti.py:115-123.
 * Everything else as gimb_loftr_forward_host. */
int gimb_loftr_host_u8_staging_bytes(int n, int ih0, int iw0, int ih1, int iw1, int h0, int w0, int h1,
                                     int w1, int with_scale, size_t* bytes);
int gimb_loftr_forward_host_u8(gimb_loftr* h, const uint8_t* img0, int ih0, int iw0, const uint8_t* img1,
                               int ih1, int iw1, const float* scale0, const float* scale1, int n, int h0,
                               int w0, int h1, int w1, void* dev_inputs, size_t dev_inputs_bytes,
                               void* workspace, size_t workspace_bytes, const gimb_loftr_out* dev_out,
                               const gimb_loftr_out* host_out, int64_t* m_out, uint64_t* h2d_bytes,
                               uint64_t* d2h_bytes, void* stream);

/* The same in two halves, for callers that overlap the upload of the next batch with the forward of the current one (what a
 * DataLoader with pin_memory + non_blocking copies does around the reference's forward): gimb_loftr_stage_host_u8 enqueues
 * the H2D copies and the device-side conversion / padding / masks of one batch on `stream` and returns WITHOUT synchronising
 * (the host buffers must stay valid and unchanged until that work has run); gimb_loftr_forward_staged_u8 runs the forward
 * from a staged buffer and reads the results back (one synchronisation at the end, like gimb_loftr_forward_host).  The caller
 * orders the two (same stream, or an event between a copy stream and the compute stream).
 * gimb_loftr_forward_host_u8 == stage + forward_staged on one stream. */
int gimb_loftr_stage_host_u8(gimb_loftr* h, const uint8_t* img0, int ih0, int iw0, const uint8_t* img1, int ih1,
                             int iw1, const float* scale0, const float* scale1, int n, int h0, int w0, int h1,
                             int w1, void* dev_inputs, size_t dev_inputs_bytes, uint64_t* h2d_bytes, void* stream);
int gimb_loftr_forward_staged_u8(gimb_loftr* h, int ih0, int iw0, int ih1, int iw1, int with_scale, int n, int h0,
                                 int w0, int h1, int w1, void* dev_inputs, size_t dev_inputs_bytes,
                                 void* workspace, size_t workspace_bytes, const gimb_loftr_out* dev_out,
                                 const gimb_loftr_out* host_out, int64_t* m_out, uint64_t* d2h_bytes, void* stream);

/* Number of kernels this library launched on behalf of handle `h` since creation (for bench.py's
 * `gpu_launches`), and the per-stage device time of the last forward when profiling is enabled. */
uint64_t gimb_loftr_launch_count(gimb_loftr* h);
/* forwards of this handle that had to repeat the coarse matching with the exact (online-max) correlation sweeps because
 * a softmax sum left the range of the fast sweeps' fixed exponent reference (csrc/corr_sweep.cu); 0 on real features */
uint64_t gimb_loftr_corr_fallbacks(gimb_loftr* h);
/* GEMM engine of the handle: 1 = tcgen05 tensor cores with split-fp16 operands (default),
 * 0 = fp32 CUDA-core kernels (the exact-fp32 cross-check path; ~10x slower). */
int gimb_loftr_set_engine(gimb_loftr* h, int engine);
/* Enable CUDA-event timing of the stages of subsequent forwards (adds host syncs; off by default). */
int gimb_loftr_set_profiling(gimb_loftr* h, int enabled);
/* names/ms arrays of length *n_stages (<= 32) for the last profiled forward. */
int gimb_loftr_last_profile(gimb_loftr* h, const char** names, float* ms, int* n_stages);

/* =================================================================================================
 * gim_dkm (DKMv3 dense matcher)
 * Replaces: networks/dkm/models/model_zoo/DKMv3.py:5-145 (construction) and
 * RegressionMatcher.match with symmetric = True, batched = False (networks/dkm/models/dkm.py:655-752).
 * The packed blob comes from gim_b200/dkm.py::pack_dkm_blob (same container layout as above).           */
typedef struct gimb_dkm gimb_dkm;

/* optional debug taps (device pointers, channels-last; index = log2(scale)): pyramid levels of pass 1, GP outputs,
 * flow [2, h/s, w/s, 2] / certainty [2, h/s, w/s] after every scale of pass 1 and of the upsample pass */
typedef struct {
  float* enc[6];
  float* gp32;
  float* gp16;
  float* flow[6];
  float* cert[6];
  float* flow_up[6];
  float* cert_up[6];
  float* dfn_flow16;  /* [2, h/16, w/16, 2] flow out of the DFN at scale 16, before its ConvRefiner */
  float* refiner_in16; /* [2, h/16, w/16, 1377] cat(x, x_hat, displacement embedding, local correlation) */
  float* refiner_dw16; /* [2, h/16, w/16, 1377] block1: depthwise 5x5 + BN + ReLU */
  float* refiner_pw16; /* [2, h/16, w/16, 1377] block1: pointwise output */
  float* refiner_out16; /* [2, h/16, w/16, 1377] after the 8 hidden blocks (input of out_conv) */
} gimb_dkm_taps;

int gimb_dkm_create(const void* blob, size_t nbytes, int device, gimb_dkm** out);
void gimb_dkm_destroy(gimb_dkm* h);
int gimb_dkm_set_engine(gimb_dkm* h, int engine);
uint64_t gimb_dkm_launch_count(gimb_dkm* h);
/* Bytes of device workspace one match() call needs.  im1 is [3, H1, W1], im2 [3, H2, W2]; (h_resized, w_resized) are the
 * caller-overwritable attributes of the reference object (trainer/lightning.py:32-37; the ZEB harness sets 660 x 880), even;
 * when upsample_preds != 0 the second pass runs at (up_h, up_w) (even).  Every stride-2 stage maps n -> ceil(n / 2) like
 * torchvision's ResNet. */
int gimb_dkm_workspace_bytes(gimb_dkm* h, int H1, int W1, int H2, int W2, int h_resized, int w_resized,
                             int upsample_preds, int up_h, int up_w, size_t* bytes);
/* match(): DEVICE fp32 NCHW images in [0, 1]; outputs warp [Hout, 2*Wout, 4] and certainty [Hout, 2*Wout] with
 * (Hout, Wout) = upsample_preds ? (up_h, up_w) : (h_resized, w_resized) - the tensors RegressionMatcher.match returns. */
/* Replaces: networks/dkm/utils/kde.py:17-26 inside RegressionMatcher.sample (dkm.py:612): Gaussian kernel density of n
 * 4-D matches, DEVICE pointers on the current device, without the n x n distance matrix. */
int gimb_kde_density(const float* points, int n, float std, float* density, void* stream);
int gimb_dkm_match(gimb_dkm* h, const float* im1, int H1, int W1, const float* im2, int H2, int W2,
                   int h_resized, int w_resized, int upsample_preds, int up_h, int up_w, void* workspace,
                   size_t workspace_bytes, float* warp, float* certainty, const gimb_dkm_taps* taps,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIMB200_H */
