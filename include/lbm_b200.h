/*
 * lbm_b200.h -- C ABI of liblbm_b200.so, the B200-native D2Q9-BGK lattice-Boltzmann solver.
 *
 * This is the drop-in boundary for the hot path of thorbenlouw/lbm-graphcore
 * (accelerate_flow -> propagate -> rebound -> collision + av_velocity).  The reference has no
 * FFI layer; its seam is the Poplar `Engine` as driven by main/LbmRunner.cpp (identically by
 * main/LbmAoS.cpp:402-424): named caller-owned host buffers plus three numbered programs and a
 * device timer.  Each entry point below names the reference call it replaces.
 *
 * Conventions (same as the reference's use of the Engine):
 *   - plain C types only; no CUDA or torch types cross the boundary;
 *   - host buffers are OWNED BY THE CALLER and are only read/written during the call;
 *   - cells are AoS float[ny][nx][9] (main/include/LatticeBoltzmannUtils.hpp:18,125-135), speeds
 *     0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE; obstacles one byte per cell, non-zero =
 *     solid, index x + y*nx (main/include/LbmParams.hpp:88-90);
 *   - every call blocks until its work is complete (engine.run is synchronous);
 *   - return 0 on success, a negative LBM_E_* code otherwise; the message for the calling thread's
 *     last failure is lbm_last_error().  Nothing throws across the ABI;
 *   - a handle is not thread-safe (the reference is single-threaded);
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     LBM_E_CUDA.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 1

enum {
  LBM_OK = 0,
  LBM_E_ARG = -1,    /* bad argument (null pointer, non-positive size, iters beyond max_iters ...) */
  LBM_E_CUDA = -2,   /* CUDA runtime/driver error, or no CUDA device                              */
  LBM_E_STATE = -3,  /* call out of order (run before upload, connect twice ...)                  */
  LBM_E_NOMEM = -4,  /* host or device allocation failed                                          */
  LBM_E_PEER = -5    /* peer access / IPC mapping between GPUs unavailable                        */
};

typedef struct lbm_handle lbm_handle;

/* ------------------------------------------------------------------------------------------------
 * The reference's operator surface (main/LbmRunner.cpp)
 * ---------------------------------------------------------------------------------------------- */

/* Engine(Executable::deserialize(exe), opts) + engine.load(device)   -- main/LbmRunner.cpp:75-96.
 * Creates a solver for a fixed problem (what the reference bakes into the compiled graph:
 * main/LbmPoplibs.cpp:304-348) on num_gpus GPUs of this process (-n/--num-ipus,
 * main/LbmRunner.cpp:19); the grid is cut into num_gpus row slabs, GPU g owning global rows
 * [g*ny/num_gpus, (g+1)*ny/num_gpus). */
int lbm_create(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega,
               int num_gpus);

/* engine.run(0): streams ">>cells" and ">>obstacles"  -- main/LbmRunner.cpp:87-88,98-100.
 * Resets the step counter; the next lbm_run starts at timestep 0. */
int lbm_upload(lbm_handle *h, const float *cells_aos, const uint8_t *obstacles);

/* engine.run(1): the whole time loop on the device -- main/LbmRunner.cpp:102-104;
 * program = accelerate_flow once + Repeat(timesteps) (main/LbmAoS.cpp:352-356).
 * Runs `iters` more timesteps (the reference always runs maxIters; splitting a run into several
 * calls gives the same result).  The first call after lbm_upload first applies the one-off guarded
 * accelerate_flow (main/LastChance.cpp:158-183, main/codelets/D2Q9Codelets.cpp:193-224). */
int lbm_run(lbm_handle *h, int iters);

/* engine.run(2): streams "<<cells" and "<<av_vel"  -- main/LbmRunner.cpp:85-86,106-108.
 * cells_aos receives the current state, av_vels the first `steps done` entries (max_iters floats
 * of room).  Either pointer may be NULL to skip that stream. */
int lbm_download(lbm_handle *h, float *cells_aos, float *av_vels);

/* engine.readTensor("readTimer") / getTileClockFrequency()  -- main/LbmRunner.cpp:134-144.
 * Device-side elapsed time of the last lbm_run (CUDA events on the launching stream; with several
 * GPUs the maximum over them). */
int lbm_elapsed_ms(lbm_handle *h, double *ms);

/* ~Engine */
void lbm_destroy(lbm_handle *h);

/* Message for the last failing call on this thread ("" if none). */
const char *lbm_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Introspection
 * ---------------------------------------------------------------------------------------------- */

int lbm_abi_version(void);
/* number of CUDA devices visible to this process (0 and LBM_E_CUDA when there is none) */
int lbm_device_count(int *count);
/* timesteps executed since the last upload */
int lbm_steps_done(lbm_handle *h, int *steps);
/* non-obstacle cells in the rows this handle owns (total_free_cells, main/LastChance.cpp:486-493) */
int lbm_free_cells(lbm_handle *h, long long *count);
/* kernel launches issued by this handle since creation (bench.py's gpu_launches) */
int lbm_launch_count(lbm_handle *h, long long *count);
/* name of the sweep kernel variant in use: "vec4" (nx % 128 == 0: 128-bit loads and stores), "hybrid" (any other
 * nx > 128: 128-bit path on the full tiles, scalar path on the ragged last tile column) or "scalar" (nx < 128;
 * LBM_B200_KERNEL=scalar forces it for comparison runs) */
const char *lbm_kernel_variant(lbm_handle *h);

/* ------------------------------------------------------------------------------------------------
 * Extensions used by the driver, the tests and bench.py (no reference counterpart)
 * ---------------------------------------------------------------------------------------------- */

/* lbm_create with an explicit device per slab.  A device may repeat: slabs that share a device run
 * back to back on one stream (how the multi-slab path is exercised on a single GPU). */
int lbm_create_on_devices(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel,
                          float omega, int num_slabs, const int *devices);

/* 2-D block partition (what the reference does across IPUs: main/include/StructuredGridUtils.hpp:501-527 cuts the grid
 * for 8 devices into 4 x 2 or 2 x 4 blocks, :805-851 gives every block a one-cell halo with wrap-around).  The grid is
 * cut into px columns x py rows of blocks; block (bx, by) -- index by * px + bx, GPU devices[index] (NULL: GPU = index; a
 * device may repeat, as in lbm_create_on_devices) -- owns columns [bx*nx/px, (bx+1)*nx/px) and rows [by*ny/py,
 * (by+1)*ny/py).  Besides the ghost rows of the row-slab partition every block keeps a west and an east ghost column and
 * four corner ghosts; the sweep kernel of the block that computes an edge cell stores its outgoing populations straight
 * into the ghost cells of the (up to eight) neighbouring blocks, over NVLink when they live on another GPU, with one
 * epoch flag per neighbour.  Everything else (upload, run, download, save) behaves as for lbm_create. */
int lbm_create_grid(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega, int px,
                    int py, const int *devices);
/* blocks per row / per column of the partition (1 x slabs for row slabs) */
int lbm_partition(lbm_handle *h, int *px, int *py);
/* like lbm_slab_info, with the block's columns as well */
int lbm_block_info(lbm_handle *h, int block, int *device, int *row_begin, int *row_count, int *col_begin, int *col_count,
                   long long *free_cells);

/* Generate the synthetic channel of SURVEY.md 8(d) on the device instead of uploading it: solid rows 0
 * and ny-1, i.i.d. solids with probability `solid_prob` elsewhere except on the driven row ny-2,
 * decided by splitmix64(seed + y*nx + x); cells get Cells::initialise values
 * (main/include/LatticeBoltzmannUtils.hpp:137-157).  Equivalent to lbm_upload of those arrays. */
int lbm_init_synthetic(lbm_handle *h, uint64_t seed, float solid_prob);

/* Partition diagnostics (what main/TileMappingStats.cpp reports for IPU tiles, here per row slab): how many slabs
 * this handle drives, and for each its GPU, its global rows and its non-obstacle cells (-1 before an upload). */
int lbm_slab_count(lbm_handle *h, int *count);
int lbm_slab_info(lbm_handle *h, int slab, int *device, int *row_begin, int *row_count, long long *free_cells);

/* the rows this handle owns: global rows [row_begin, row_begin + row_count) */
int lbm_rows(lbm_handle *h, int *row_begin, int *row_count);

/* per-timestep sum of sqrtf(u_sq) over this handle's non-obstacle cells, in double (one value per
 * step done).  av_vels[t] = (float)(sum over ranks, in rank order) / (float)(total free cells). */
int lbm_download_tot_u(lbm_handle *h, double *tot_u);

/* The four computed columns of final_state.dat -- u_x, u_y, u, pressure per cell, float[rows][nx][4] -- evaluated
 * on the device from the current state with the arithmetic of lbm::writeResults
 * (main/include/LatticeBoltzmannUtils.hpp:240-267): 16 bytes per cell cross PCIe instead of the 36-byte cell, and
 * the host only has to format text.  Rows as for lbm_download. */
int lbm_download_fields(lbm_handle *h, float *fields);

/* Prepare once, run many.  The reference compiles and serialises its solver once (main/LbmPoplibs.cpp:375-389) and
 * lbm_runner deserialises it with --exe (main/LbmRunner.cpp:20,75-84).  Here lbm_save writes the prepared problem and
 * its current state -- parameters, obstacles, lattice, step counter and the per-step sums so far -- to `path`, and
 * lbm_load creates a solver from such a file on num_gpus GPUs (any number: the file holds host interchange formats,
 * not device layouts).  A loaded solver continues exactly where the saved one stopped: lattice bit-identical, av_vels
 * of all steps identical to an uninterrupted run on the same number of GPUs.  Not available to a rank-mode handle
 * that owns only part of the grid.  (Within one handle, the time loop of small grids is also kept as an instantiated
 * CUDA graph and replayed by every later lbm_run: LBM_B200_GRAPH=0 disables, =1 forces it for any size.) */
int lbm_save(lbm_handle *h, const char *path);
int lbm_load(lbm_handle **out, const char *path, int num_gpus);
/* the same onto a px x py block partition (arguments as for lbm_create_grid) */
int lbm_load_grid(lbm_handle **out, const char *path, int px, int py, const int *devices);
/* the problem a handle was created (or loaded) for; any pointer may be NULL */
int lbm_problem(lbm_handle *h, int *nx, int *ny, int *max_iters, float *density, float *accel, float *omega);

/* One whole job from host buffers -- what main/LbmRunner.cpp:98-108 does with run(0), run(1), run(2) -- as ONE call that
 * overlaps the three: the grid is solved in horizontal bands (a band plus a margin of >= iters rows either side is all
 * that `iters` timesteps of the band depend on), so that uploads, sweeps and downloads of different bands run concurrently
 * and both directions of PCIe are busy at once.  Like the Engine, the pipeline is prepared once for a fixed problem
 * (lbm_pipeline_create allocates its band solvers) and can then solve any number of inputs of that shape:
 *   cells_in / obstacles / cells_out / av_vels as for lbm_upload / lbm_download; cells_out must NOT alias cells_in (the
 *   margins of a band are read again after its neighbours have been written); pinned buffers (lbm_host_alloc) for full speed.
 * Same results as upload + run + download on one GPU: lattice bit-identical, av_vels from per-band sums added in band
 * order.  Grids too small (or step counts too large) for bands to pay run the three calls on one internal handle.
 * lbm_solve_host = create + solve + destroy. */
typedef struct lbm_pipeline lbm_pipeline;
int lbm_pipeline_create(lbm_pipeline **out, int nx, int ny, int iters, float density, float accel, float omega, int device);
int lbm_pipeline_solve(lbm_pipeline *p, const float *cells_in, const uint8_t *obstacles, float *cells_out, float *av_vels);
/* The same for ONE RANK of a multi-process job: the pipeline solves the WINDOW of global rows [row_begin, row_begin +
 * row_count) (row_count a multiple of 8) of the nx x ny grid.  cells_window / obstacles_window hold the initial state of the
 * window AND of `margin_rows` (lbm_pipeline_info) rows below and above it, contiguously, in global row order with periodic
 * wrap -- (row_count + 2 * margin_rows) rows; the rows outside the window belong to the neighbouring ranks, which exchange
 * them once (host to host) before the job.  In return no rank waits for another during the job: `iters` timesteps of the
 * window depend on nothing else.  cells_out receives the window's row_count rows; tot_u[iters] and *free_cells are this
 * rank's part of the av_vels sums, to be combined in rank order like lbm_download_tot_u's.  A band plus its two margins is
 * never taller than the grid (it would hold the driven row twice): short grids get shorter bands, and LBM_E_ARG when not even
 * one 8-row band fits (ny < 8 + 2 * margin_rows). */
int lbm_pipeline_create_window(lbm_pipeline **out, int nx, int ny, int row_begin, int row_count, int iters, float density,
                               float accel, float omega, int device);
int lbm_pipeline_solve_window(lbm_pipeline *p, const float *cells_window, const uint8_t *obstacles_window, float *cells_out,
                              double *tot_u, long long *free_cells);
/* how a job WOULD be cut, without creating anything (no GPU needed): window_rows = 0 for the whole grid, else one rank's
 * row count.  bands = 1 and margin_rows = 0 mean "not banded". */
int lbm_pipeline_plan(int nx, int ny, int window_rows, int iters, int *bands, int *band_rows, int *margin_rows);
/* how the job is cut (1 band = not banded) and the wall time of the last solve; any pointer may be NULL */
int lbm_pipeline_info(lbm_pipeline *p, int *bands, int *band_rows, int *margin_rows, double *last_ms);
void lbm_pipeline_destroy(lbm_pipeline *p);
int lbm_solve_host(int nx, int ny, int iters, float density, float accel, float omega, int device, const float *cells_in,
                   const uint8_t *obstacles, float *cells_out, float *av_vels, double *elapsed_ms);

/* Self-test of the kernels' IEEE division / square-root helpers against CUDA's __fdiv_rn / __fsqrt_rn on
 * GPU `device`: exhaustive for x/9, x/36 and sqrt(x), `pairs` random operand pairs for a/b.
 * counts[7] = {x/9 mismatches, sqrt mismatches, a/b mismatches, x/36 mismatches, #x/9, #sqrt, #a/b}. */
int lbm_selftest_math(int device, unsigned long long pairs, unsigned long long *counts);

/* pinned host memory for callers that want full-speed uploads/downloads */
void *lbm_host_alloc(size_t bytes);
void lbm_host_free(void *p);
/* the same, placed on the NUMA node of GPU `device`'s PCIe root when the OS allows it (free with lbm_host_free) */
void *lbm_host_alloc_near(size_t bytes, int device);
/* restrict the calling thread to the CPUs local to GPU `device`; returns the NUMA node (>= 0) or a negative code when
 * the topology is unknown or the call is refused.  Optimisation only. */
int lbm_bind_thread_near(int device);

/* ------------------------------------------------------------------------------------------------
 * One process per GPU (torchrun): each rank owns one row slab; neighbours' lattices are mapped with
 * CUDA IPC and the sweep kernel stores its outgoing halo rows straight into peer memory over
 * NVLink.  lbm_upload/lbm_download then take/return only the owned rows: float[row_count][nx][9] and
 * uint8[row_count][nx].  Protocol (B = a barrier across all ranks, the caller's job -- the library has no
 * communicator of its own):
 *
 *   create_rank -> ipc_export -> (all-gather the blobs) -> ipc_connect -> B
 *   upload | init_synthetic                 resets this rank's halo flags
 *   B                                       every rank's lattice and flags are in place before anyone pushes a halo
 *   run ... run, download                   no barrier needed between these
 *   B -> upload | init_synthetic -> B       RE-UPLOAD: the barrier BEFORE it guarantees that no neighbour is still
 *                                           pushing halo rows or epochs of the previous job into the slab being reset
 *   B -> ipc_disconnect -> B -> destroy     TEARDOWN: all ranks finished running; every rank closes its mappings of
 *                                           the neighbours' slabs; only then may a slab be freed (freeing memory that
 *                                           another process still has mapped is undefined behaviour in CUDA)
 *
 * lbm-graphcore_b200/slabs.py (connect_ring, upload barrier helpers, close_ring) implements this over
 * torch.distributed.
 * ---------------------------------------------------------------------------------------------- */

int lbm_create_rank(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel,
                    float omega, int rank, int world, int device);
/* the same for a px x py BLOCK partition (lbm_create_grid's layout, one rank per block; rank = by * px + bx): up to eight
 * neighbouring ranks are mapped by lbm_ipc_connect.  lbm_upload / lbm_download / lbm_download_fields then take / return only
 * the rank's block, row_count rows of col_count cells (lbm_block_info(h, 0, ...)). */
int lbm_create_rank_grid(lbm_handle **out, int nx, int ny, int max_iters, float density, float accel, float omega, int rank,
                         int px, int py, int device);
size_t lbm_ipc_blob_bytes(void);
int lbm_ipc_export(lbm_handle *h, void *blob);
/* blobs: world * lbm_ipc_blob_bytes() bytes, rank-major */
int lbm_ipc_connect(lbm_handle *h, const void *blobs);
/* waits for this rank's queued work, then closes its mappings of the neighbours' slabs (first half of the
 * teardown above).  Afterwards the handle can only be downloaded from and destroyed. */
int lbm_ipc_disconnect(lbm_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
