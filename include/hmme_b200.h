/*
 * hmme_b200.h -- C ABI of libhmme_b200.so: whole-CTU integer-pel block-matching motion estimation
 * for HM-OpenCL's TEncOpenCL path, hand-written CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary.  In the reference the "FFI" is the public
 * surface of a compiled-in C++ class, /root/reference/source/Lib/TLibEncoder/TEncOpenCL.h:105-123;
 * the replacement class (hm-opencl_b200/host/TEncOpenCL.{h,cpp}) keeps those signatures and calls
 * the functions below.  Each entry point names the reference interface it replaces.
 *
 * Conventions: every function returns 0 on success and a negative hmme_status on failure;
 * hmme_last_error() gives the message.  There is NO CPU fallback: without a usable sm_100 device
 * hmme_create fails.  All calls are synchronous unless the name ends in _async; a context is used
 * from one host thread at a time (the encoder is single-threaded, SURVEY.md section 8b "Threading").
 *
 * Result layout: 593 entries per (CTU, reference picture) job in the order of
 * TComDataCU::getIndexBlock (TComDataCU.cpp:4676-6461): X/Y = integer-pel MV of the winner,
 * sad = pure SAD at the winner (TEncOpenCL::getRuiCost), cost = SAD + MV-bit cost (minSad).
 */
#ifndef HMME_B200_H
#define HMME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMME_NUM_CTU_PARTS 593 /* TypeDef.h:263 (AMP_ENC_SPEEDUP = 0) */
#define HMME_CTU_SIZE 64       /* TEncSearch.cpp:3745 hard-codes the 64x64 2Nx2N trigger */

typedef enum {
    HMME_OK = 0,
    HMME_ERR_ARG = -1,        /* bad pointer / size / range */
    HMME_ERR_NO_DEVICE = -2,  /* no CUDA device, or not compute capability 10.x */
    HMME_ERR_CUDA = -3,       /* a CUDA runtime call failed */
    HMME_ERR_RANGE = -4,      /* search range / CTU size beyond what the context was created for */
    HMME_ERR_CONTENT = -5,    /* plane declared 8-bit holds samples outside [0,255] */
    HMME_ERR_BOUNDS = -6,     /* a job's window leaves the plane allocation (reference: undefined behaviour, App. B4) */
    HMME_ERR_NCCL = -7        /* NCCL could not be loaded or a collective failed (only HMME_REF_BROADCAST needs NCCL) */
} hmme_status;

typedef struct hmme_ctx hmme_ctx;

/* One search job: the CTU at picture position (ctuX, ctuY) searched over candidates
 * (ltx + x, lty + y), x,y in 0..2R (pcMvSrchRngLT of TEncOpenCL::calcMotionVectors). */
typedef struct { int32_t ctuX, ctuY, ltx, lty; } hmme_job;

/* A luma plane resident in device memory.  `base` is the FIRST byte of the allocation (top-left of
 * the margin); picture sample (0,0) is at base + (marginY*pitch + marginX)*elemBytes.  elemBytes is
 * 1 (8-bit samples, the fast path) or 2 (int16, e.g. the bi-prediction "current" block 2*org-pred,
 * TEncSearch.cpp:3702-3712).  pitch is in elements.  Rows 0..height+2*marginY-1 are addressable; base is 16-byte
 * aligned and the allocation extends 64 bytes past the last row (hmme_plane_alloc does both; memory wrapped from elsewhere,
 * e.g. a torch tensor, must too: the kernels fetch window rows with 16-byte granular TMA bulk copies). */
typedef struct {
    void* base;
    int32_t elemBytes, pitch, width, height, marginX, marginY;
} hmme_plane;

/* ---- discovery / lifetime: TEncOpenCL::findDevice, compileKernelSource, createBuffers (TEncOpenCL.cpp:69-238) */
int hmme_device_count(int* count);
int hmme_create(hmme_ctx** out, int device, int maxCtuW, int maxCtuH, int maxSearchRange);
void hmme_destroy(hmme_ctx* ctx);
const char* hmme_device_name(hmme_ctx* ctx);          /* TEncOpenCL::getDeviceInfo */
const char* hmme_last_error(hmme_ctx* ctx);           /* ctx may be NULL: error of the last failed hmme_create */
void* hmme_stream(hmme_ctx* ctx);                     /* the cudaStream_t all work of this context runs on */
void* hmme_host_alloc(size_t bytes);                  /* page-locked host memory (asynchronous copies are only asynchronous from / to it) */
void hmme_host_free(void* p);

/* ---- lambda: TEncOpenCL::setLambda (TEncOpenCL.h:121), m_lambda = (UInt)floor(65536*sqrt(lambda)) */
int hmme_set_lambda(hmme_ctx* ctx, double lambda);
int hmme_set_lambda_q16(hmme_ctx* ctx, uint32_t lambdaQ16);
uint32_t hmme_get_lambda_q16(hmme_ctx* ctx);

/* ---- synchronous per-CTU search: TEncOpenCL::calcMotionVectors + getX/getY/getRuiCost
 * (TEncOpenCL.cpp:240-362, TEncSearch.cpp:3749-3764).  Host pointers, borrowed for the call.
 *   cur      : 64x64 int16, stride curStride
 *   refAtCtu : pointer into the padded int16 reference plane at the CTU origin, stride refStride;
 *              samples [lt, lt + 2R + 63] in both axes are read with LINEAR addressing (App. B4)
 * 8-bit content in both takes the packed-SAD kernel, anything else the exact 16-bit kernel.
 * cost may be NULL. */
int hmme_search_ctu(hmme_ctx* ctx, const int16_t* cur, int curStride, const int16_t* refAtCtu, int refStride,
                    int range, int ltx, int lty, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);

/* ---- device-resident planes (whole-frame batching; reference pictures are uploaded once per picture,
 * the hook being where TComSlice::setRefPicList extends the borders, TComSlice.cpp:351-377) */
int hmme_plane_alloc(hmme_ctx* ctx, hmme_plane* out, int elemBytes, int width, int height, int marginX, int marginY);
int hmme_plane_free(hmme_ctx* ctx, hmme_plane* plane);
/* Host int16 plane (HM's Pel) -> device plane; narrows to 8 bit on the device when plane->elemBytes == 1
 * and fails with HMME_ERR_CONTENT if a sample does not fit.  hostOrigin points at picture sample (0,0);
 * the margins are copied too (they must exist on the host side, as in TComPicYuv, TComPicYuv.cpp:93-94). */
int hmme_plane_upload_s16(hmme_ctx* ctx, const hmme_plane* plane, const int16_t* hostOrigin, int hostStride);
/* Host 8-bit plane -> device 8-bit plane (same geometry rules): half the bytes of the int16 form for callers that hold 8-bit video. */
int hmme_plane_upload_u8(hmme_ctx* ctx, const hmme_plane* plane, const uint8_t* hostOrigin, int hostStride);
int hmme_plane_upload_u8_async(hmme_ctx* ctx, const hmme_plane* plane, const uint8_t* hostOrigin, int hostStride);
/* The general form: only the rectangle [x0, x1) x [y0, y1) (picture coordinates; negative / beyond the picture = margin samples) of
 * a host plane whose samples are hostElemBytes wide (1: uint8, 2: int16).  This is how a GPU that owns a band of CTU rows receives
 * just the rows it reads (band + search-window halo).  Asynchronous, same rules as hmme_plane_upload_s16_async. */
int hmme_plane_upload_rect_async(hmme_ctx* ctx, const hmme_plane* plane, const void* hostOrigin, int hostStride, int hostElemBytes,
                                 int x0, int y0, int x1, int y1);

/* ---- whole-frame batch: njobs independent calcMotionVectors calls in one launch sequence.
 * jobs, X, Y, sad, cost are HOST arrays ([njobs] and [njobs][593]); cost may be NULL.
 * Enqueues H2D(jobs) -> kernels -> D2H(results) on the context stream and waits. */
int hmme_search_frame(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs,
                      int range, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);
/* Same, but only enqueues the kernels (jobs already copied by this call, results left in the context's
 * device result buffer): the kernel-only leg of bench.py.  Pair with hmme_fetch_results. */
int hmme_search_frame_async(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs,
                            int njobs, int range);
int hmme_fetch_results(hmme_ctx* ctx, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);
/* Thread blocks the search of njobs jobs at +-range launches, and how many the device runs at once (one per SM): the length of a
 * frame in waves, which decides how many frames a pipelining caller keeps in flight (hmme_group_pipeline_depth). */
int hmme_search_launch_size(hmme_ctx* ctx, int njobs, int range, int* ctas, int* ctasPerWave);
/* Fully asynchronous legs for pipelining frames over two contexts (copies of one context overlap the kernels of the
 * other): enqueue only; host buffers must stay valid (and should be page-locked) until hmme_sync returns.  The 8-bit
 * content check of an asynchronous upload is reported by the next hmme_sync / synchronous call (HMME_ERR_CONTENT). */
int hmme_plane_upload_s16_async(hmme_ctx* ctx, const hmme_plane* plane, const int16_t* hostOrigin, int hostStride);
int hmme_fetch_results_async(hmme_ctx* ctx, int njobs, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);
int hmme_sync(hmme_ctx* ctx);

/* ---- device-resident result tables.  The reference keeps TComMv allMotionVectors[2][33][593] and
 * Distortion allRuiCost[2][33][593] (TEncSearch.h:114-115) for the CTU being coded, on the host.  A table keeps the same 593-entry
 * records for EVERY CTU job of a picture in HBM, one slot per (reference list, reference index) -- or per window hypothesis of the
 * speculative whole-frame search -- so that a B picture's two lists do not overwrite each other.
 * hmme_search_frame_table_async leaves its results in table[slot] (nothing is copied to the host); hmme_table_fetch_async copies a
 * job range of a slot to host arrays; hmme_table_device_ptr gives the device address of one of the four [jobsPerSlot][593] arrays
 * (0: X, 1: Y, 2: sad, 3: cost) for consumers on the GPU. */
typedef struct hmme_table hmme_table;
int hmme_table_create(hmme_ctx* ctx, hmme_table** out, int slots, int jobsPerSlot);
void hmme_table_destroy(hmme_table* table);
int hmme_search_frame_table_async(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, const hmme_job* jobs, int njobs,
                                  int range, hmme_table* table, int slot);
int hmme_table_fetch_async(hmme_ctx* ctx, hmme_table* table, int slot, int firstJob, int njobs, int32_t* X, int32_t* Y,
                           uint32_t* sad, uint32_t* cost);
const void* hmme_table_device_ptr(hmme_table* table, int slot, int array);

/* ---- multi-GPU (the reference drives one device, TEncOpenCL.cpp:129,185 -- this is north_star's addition).
 * A group searches ONE frame on several B200s: the job list is cut into contiguous bands (CTU rows, cut at CTU granularity, sizes
 * differ by at most one job), every GPU searches its band and writes its rows of the caller's [njobs][593] host tables.
 * Reference-picture distribution:
 *   HMME_REF_BAND_HALO  every GPU copies, over its own PCIe link, only the rectangle its band reads (band + window halo); no collective
 *   HMME_REF_BROADCAST  global rank 0 uploads the whole padded plane, ncclBroadcast carries it over NVLink / NVSwitch
 * Process models: hmme_group_create = one process, ndev GPUs (one enqueue thread per GPU; ncclCommInitAll);
 * hmme_group_create_rank = this process is rank `rank` of `nranks` one-GPU processes (torchrun); uniqueId (128 bytes, made by rank 0
 * with hmme_group_unique_id and distributed by the launcher) may be NULL when only HMME_REF_BAND_HALO is used.  With create_rank every
 * process passes the same whole-frame arguments and fills only its own band's rows of the tables.
 * Host planes: hostElemBytes 2 = HM's Pel (int16, narrowed on the device), 1 = uint8.  Origins point at picture sample (0,0); the
 * reference plane's margins must exist (TComPicYuv).  Windows must stay inside the padded picture (HMME_ERR_BOUNDS otherwise).
 * slot (0 .. HMME_GROUP_SLOTS - 1) selects one of three frames in flight; host buffers stay valid (ideally page-locked) until
 * hmme_group_sync(slot).  Callers that pipeline frames cycle through hmme_group_pipeline_depth() slots (two are enough while a GPU's band
 * is large; narrow bands on many GPUs need the third to keep a search queued while a slot is being refilled). */
typedef struct hmme_group hmme_group;
enum { HMME_REF_BAND_HALO = 0, HMME_REF_BROADCAST = 1 };
#define HMME_GROUP_SLOTS 3
int hmme_group_create(hmme_group** out, const int* devices, int ndev, int maxSearchRange);
int hmme_group_unique_id(void* id, size_t bytes);
int hmme_group_create_rank(hmme_group** out, int device, int rank, int nranks, const void* uniqueId, int maxSearchRange);
void hmme_group_destroy(hmme_group* group);
const char* hmme_group_last_error(hmme_group* group);            /* group may be NULL: error of the last failed create */
int hmme_group_size(hmme_group* group, int* world, int* nlocal);
int hmme_group_set_lambda_q16(hmme_group* group, uint32_t lambdaQ16);
int hmme_group_configure(hmme_group* group, int width, int height, int marginX, int marginY, int refDist);
int hmme_group_search_frame_async(hmme_group* group, int slot, const void* curHostOrigin, int curHostStride,
                                  const void* refHostOrigin, int refHostStride, int hostElemBytes, const hmme_job* jobs, int njobs,
                                  int range, int32_t* X, int32_t* Y, uint32_t* sad, uint32_t* cost);
int hmme_group_sync(hmme_group* group, int slot);                /* slot -1: all */
int hmme_group_search_frame(hmme_group* group, const void* curHostOrigin, int curHostStride, const void* refHostOrigin,
                            int refHostStride, int hostElemBytes, const hmme_job* jobs, int njobs, int range, int32_t* X, int32_t* Y,
                            uint32_t* sad, uint32_t* cost);
/* How many frames a pipelining caller should keep in flight (slots to cycle through) for frames of njobs jobs at +-range: 2 or 3. */
int hmme_group_pipeline_depth(hmme_group* group, int njobs, int range);
int hmme_group_band(hmme_group* group, int njobs, int localIndex, int* first, int* count);
hmme_ctx* hmme_group_context(hmme_group* group, int localIndex, int slot);   /* the per-GPU context behind a slot (refinement, timing) */
int hmme_group_last_kernel_ms(hmme_group* group, int slot, float* searchKernelMsMaxOverLocalGpus);
uint64_t hmme_group_kernel_launches(hmme_group* group);
/* The band arithmetic itself, pure host code: jobs [first, first + count) of njobs belong to `rank` of `world`; and the picture
 * rectangles {x0, y0, x1, y1} a job range reads from the current frame (its CTUs) and from the reference picture (band + halo). */
int hmme_band_split(int njobs, int world, int rank, int* first, int* count);
int hmme_band_extent(const hmme_job* jobs, int njobs, int range, int* curRect, int* refRect);

/* ---- CUDA graphs for launch-bound steps (many GPUs, narrow bands: tens of runtime calls for a fraction of a millisecond of
 * kernels).  hmme_graph_begin .. hmme_graph_end records the asynchronous calls made on this context in between
 * (hmme_plane_upload_s16_async, hmme_search_frame_async, hmme_fetch_results_async, hmme_refine_frame_async,
 * hmme_fetch_frac_async) instead of running them; hmme_graph_launch replays them with one call on the context's stream.
 * Host buffers must be page-locked and keep their addresses (their CONTENT is read / written at every launch); run the step
 * once before capturing so that every device buffer has its final size.  Synchronous calls are not allowed while capturing. */
typedef struct hmme_graph hmme_graph;
int hmme_graph_begin(hmme_ctx* ctx);
int hmme_graph_end(hmme_ctx* ctx, hmme_graph** out);
int hmme_graph_launch(hmme_ctx* ctx, hmme_graph* graph);
void hmme_graph_destroy(hmme_graph* graph);

/* ---- fractional-pel refinement, the step right after the integer search:
 * TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:4294-4331) = half-pel then quarter-pel refinement (xPatternRefinement,
 * :816-872) over HEVC 8-tap interpolated samples (xExtDIFUpSamplingH/Q, :5386-5600; TComInterpolationFilter.cpp:155-250) with
 * the Hadamard (TComRdCost::xGetHADs, TComRdCost.cpp:1537-1600) or SAD distortion plus lambda*bits(mv - predictor)>>16
 * (TComRdCost.h:166-185).  lambda is the context's (hmme_set_lambda: m_uiLambdaMotionSAD has the same quantisation).
 * A PU: luma rectangle (multiples of 4, up to 64), the integer-pel MV to refine and the quarter-pel predictor
 * (TComRdCost::setPredictor).  cur may be an int16 plane (bi-prediction target), ref must be 8-bit.  The reference plane
 * needs 4 samples (+ up to 4 of tile padding) around every MV-displaced PU.  useHad = HadamardME && !lossless. */
typedef struct { int32_t x, y, w, h, mvx, mvy, predx, predy; } hmme_pu;
/* final quarter-pel MV (4*integer + 2*half + quarter), ruiCost as xPatternSearchFracDIF returns it, and cost minus its MV part */
typedef struct { int32_t mvx, mvy; uint32_t cost, dist; } hmme_frac_result;
/* Host PU list -> results[npus]; candCosts is NULL or [npus][18]: the cost of each half- then quarter-pel candidate in the
 * order of the reference's tables (TEncSearch.cpp:51-75), for tests.  Synchronous. */
int hmme_refine_frac(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, const hmme_pu* pus, int npus, int useHad,
                     hmme_frac_result* results, uint32_t* candCosts);
/* One PU with HOST pointers, synchronous -- what the body of xPatternSearchFracDIF needs: cur = the pattern key block (w x h,
 * int16, stride curStride; original samples or the bi-prediction target 2*org - pred of 8-bit video, i.e. values in [-255, 510] --
 * the Hadamard intermediates are packed into 16 bits, exact for |cur - prediction| <= 4095), refAtPu = piRefY, i.e. a pointer INTO the padded int16 reference plane at
 * the PU origin (samples [mv-4, mv+w+3] x [mv-4, mv+h+3] around it are read), integer MV, quarter-pel predictor.  Returns the
 * final quarter-pel MV (4*mv + 2*half + quarter), ruiCost, and optionally cost minus the MV cost. */
int hmme_refine_pu(hmme_ctx* ctx, const int16_t* cur, int curStride, const int16_t* refAtPu, int refStride, int w, int h,
                   int mvx, int mvy, int predx, int predy, int useHad, int32_t* mvQpelX, int32_t* mvQpelY, uint32_t* cost,
                   uint32_t* dist);
/* All 593 partitions of every job of the preceding hmme_search_frame[_async] on this context, starting from that search's
 * integer winners, which never leave the device.  predsQpel: NULL (zero predictor) or [njobs][2] quarter-pel predictors.
 * results: [njobs][593].  The _async pair only enqueues (pair with hmme_sync). */
int hmme_refine_frame(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, int njobs, const int32_t* predsQpel,
                      int useHad, hmme_frac_result* results);
int hmme_refine_frame_async(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, int njobs, const int32_t* predsQpel,
                            int useHad);
int hmme_fetch_frac_async(hmme_ctx* ctx, int njobs, hmme_frac_result* results);
int hmme_last_frac_ms(hmme_ctx* ctx, float* refineKernelMs);   /* most recent refinement or hmme_mc_cost kernel */

/* ---- distortion of the motion-compensated uni-prediction of PUs at given QUARTER-PEL MVs: the
 * arithmetic inside TEncSearch::xGetTemplateCost (TEncSearch.cpp:3634-3674: xPredInterBlk + SAD of each AMVP candidate; the
 * floating-point calcRdCost on top stays with the caller) and of the uni-directional candidates of xMergeEstimation /
 * xGetInterPredictionError (:2814-2836, Hadamard when HadamardME).  The MV must already be clipped (TComDataCU::clipMv).
 * dist[i] = SAD (useHad = 0) or Hadamard SATD (xGetHADs) between the current block and the prediction.  Same plane rules as
 * hmme_refine_frac, around the integer part of the MV.  Synchronous. */
typedef struct { int32_t x, y, w, h, mvQpelX, mvQpelY; } hmme_mc_pu;
int hmme_mc_cost(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref, const hmme_mc_pu* pus, int npus, int useHad,
                 uint32_t* dist);
/* Bi-directional PUs (merge candidates / the motion-estimation result that xGetInterPredictionError evaluates through
 * TComPrediction::xPredInterBi, TComPrediction.cpp:603-651): each list contributes its 14-bit intermediate, the prediction is
 * TComYuv::addAvg's clip((p0 + p1 + offset) >> 7) (TComYuv.cpp:352-410).  ref0 / ref1 are the two reference planes, the MVs are
 * clipped quarter-pel MVs.  Identical motion (same picture, same MV) is the caller's case to route to hmme_mc_cost
 * (TComPrediction::xCheckIdenticalMotion, :501-516). */
typedef struct { int32_t x, y, w, h, mv0QpelX, mv0QpelY, mv1QpelX, mv1QpelY; } hmme_mc_bi_pu;
int hmme_mc_cost_bi(hmme_ctx* ctx, const hmme_plane* cur, const hmme_plane* ref0, const hmme_plane* ref1, const hmme_mc_bi_pu* pus,
                    int npus, int useHad, uint32_t* dist);
int hmme_mc_cost_bi_pu(hmme_ctx* ctx, const int16_t* cur, int curStride, const int16_t* ref0AtPu, int ref0Stride, int mv0QpelX,
                       int mv0QpelY, const int16_t* ref1AtPu, int ref1Stride, int mv1QpelX, int mv1QpelY, int w, int h, int useHad,
                       uint32_t* dist);
/* One PU with HOST pointers, synchronous (the arguments xGetTemplateCost has): cur = original block (int16, stride curStride),
 * refAtPu = pointer into the padded int16 reference plane at the PU origin, clipped quarter-pel MV. */
int hmme_mc_cost_pu(hmme_ctx* ctx, const int16_t* cur, int curStride, const int16_t* refAtPu, int refStride, int w, int h,
                    int mvQpelX, int mvQpelY, int useHad, uint32_t* dist);

/* ---- measurement hooks (bench.py / profiles): CUDA-event time of the dominant kernel of the most recent
 * search call on this context's stream, kernel launches issued so far, and the integer-ALU issue-rate
 * micro-benchmark that fixes the roofline denominator. */
int hmme_last_kernel_ms(hmme_ctx* ctx, float* searchKernelMs);
uint64_t hmme_kernel_launches(hmme_ctx* ctx);
int hmme_measure_int_alu_peak(hmme_ctx* ctx, double* laneOpsPerSec, double* lanesPerClkPerSm, double* smMhz);

/* ---- the 593-entry layout, for callers that want it without linking HM (index -> x, y, w, h) */
int hmme_partition_rect(int index, int* x, int* y, int* w, int* h);
/* Search-window placement: TEncSearch::xSetSearchRange (TEncSearch.cpp:3814-3830) + TComDataCU::clipMv (TComDataCU.cpp:2907-2920)
 * for a 64x64 CTU at (cuX, cuY): quarter-pel centre MV -> integer-pel left/top (what calcMotionVectors receives as
 * pcMvSrchRngLT, i.e. a job's ltx/lty) and right/bottom (unused by the GPU path, may be NULL). */
int hmme_search_window(int predHorQpel, int predVerQpel, int range, int cuX, int cuY, int picWidth, int picHeight,
                       int* ltx, int* lty, int* rbx, int* rby);
/* Closed form of TComDataCU::getIndexBlock (TComDataCU.cpp:3379-6464): PartSize enum value, CU depth, PU index, z-order
 * index of the CU in 4x4 units, CU width/height -> index 0..592, or -1 exactly where the reference's switch has no case. */
int hmme_index_block(int partSize, int depth, int partIdx, int absZIdxInCtu, int cuWidth, int cuHeight);
const char* hmme_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HMME_B200_H */
