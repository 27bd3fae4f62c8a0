/* b200fft -- C ABI of the B200-native distributed R2C FFT engine (libb200fft.so).
 *
 * Drop-in boundary for mpiFFT4py's hot path.  The reference has no FFI of its own: its compiled
 * boundary is the duck-typed serial-FFT function table (mpiFFT4py/serialFFT/__init__.py:1-6,
 * pyfftw_fft.py:26-203, numpy_fft.py:25-107), the Cython helpers (mpiFFT4py/cython/maths.pyx:9-43)
 * and mpi4py collectives called from slab.py / pencil.py / line.py.  Each entry point below names
 * the reference interface it replaces.  Plain pointers and sizes only; every data pointer is a
 * DEVICE pointer of the current CUDA device unless stated otherwise; `stream` is a cudaStream_t
 * passed as void*.  All functions return 0 on success, a non-zero code otherwise
 * (b200fft_last_error() gives the message of the calling thread's last failure).
 */
#ifndef B200FFT_H
#define B200FFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200FFT_API __attribute__((visibility("default")))
#else
#define B200FFT_API
#endif

#define B200FFT_MAXP 16 /* max ranks of one exchange (one NVSwitch box has 8 GPUs) */

enum { B200FFT_SINGLE = 0, B200FFT_DOUBLE = 1 };              /* mpibase.py:133-137 datatypes() */
enum { B200FFT_SLAB = 0, B200FFT_PENCIL_X = 1, B200FFT_PENCIL_Y = 2, B200FFT_LINE = 3,
       B200FFT_SLAB_C2C = 4 /* slab.C2C, slab.py:538-825: u and fu are both complex */ };
enum { B200FFT_DEALIAS_NONE = 0, B200FFT_DEALIAS_3_2 = 1, B200FFT_DEALIAS_2_3 = 2 };
enum { B200FFT_PIPELINE_AUTO = 0, B200FFT_PIPELINE_X = 1, B200FFT_PIPELINE_KZ = 2 };
enum { B200FFT_LAYOUT_YBLOCK = 0, B200FFT_LAYOUT_NATURAL = 1 };
enum { B200FFT_TRANSPORT_NCCL = 0, B200FFT_TRANSPORT_P2P = 1,
       B200FFT_TRANSPORT_STORE = 2 /* fused: the producing FFT pass stores into the peers' buffers */ };

enum {
  B200FFT_OK = 0,
  B200FFT_ERR_ARG = 1,       /* bad argument (AssertionError in the reference) */
  B200FFT_ERR_RANKS = 2,     /* illegal rank count (IOError: slab.py:89-91, pencil.py:201-205) */
  B200FFT_ERR_UNSUPPORTED = 3, /* length without a radix plan / too long for shared memory */
  B200FFT_ERR_CUDA = 4,
  B200FFT_ERR_NCCL = 5,
  B200FFT_ERR_NOMEM = 6
};

B200FFT_API int b200fft_version(void);
B200FFT_API const char* b200fft_last_error(void);
/* 1 if complex length n has a kernel plan (n = 2^k or 3*2^k within the supported range) */
B200FFT_API int b200fft_supported_length(int n);

/* ---------------------------------------------------------------------------------------------
 * Low level: one fused FFT pass.  These are what serialFFT.fft/ifft/rfft/irfft (and the copies
 * around them) become.  Element offsets, not bytes.
 * ------------------------------------------------------------------------------------------- */

/* One side (load or store) of a pass: the transformed-axis index i (after pad/truncate mapping to
 * the physical extent nphys) is cut into nchunk chunks of `chunk` entries, the last one taking
 * the remainder; chunk p lives at base[p]:  base[p] + b*sb[p] + (i - p*chunk)*si[p] + j.
 * nchunk == 1 is a plain strided array; nchunk > 1 is the per-peer block layout of an exchange --
 * the Alltoallw subarray datatypes of slab.py:199-211 / pencil.py:218-246,971-999 and the
 * rollaxis / transpose_Uc packs (maths.pyx:21-31, pencil.py:109-143, line.py:14-24). */
typedef struct {
  void* base[B200FFT_MAXP];
  long long sb[B200FFT_MAXP];
  long long si[B200FFT_MAXP];
  int chunk;
  int nchunk;
  int nphys; /* physical extent along the FFT axis; < n means zero-pad (load) / truncate (store) */
} b200fft_side_t;

/* 2/3-rule mask folded into a load (dealias_filter maths.pyx:9-19; get_dealias_filter
 * slab.py:191-197, pencil.py:343-349, line.py:131-136).  Zero when any band contains the index:
 * i+i_off in [i_lo,i_hi]; b+b_off in [b_lo,b_hi]; j/jdiv+jq_off in [jq_lo,jq_hi];
 * j%jdiv+jr_off in [jr_lo,jr_hi].  Disabled bands use lo > hi. */
typedef struct {
  int on;
  int i_off, i_lo, i_hi;
  int b_off, b_lo, b_hi;
  int jdiv;
  int jq_off, jq_lo, jq_hi;
  int jr_off, jr_lo, jr_hi;
} b200fft_mask_t;

/* Batched complex FFT of length n along the strided middle axis of [B][n][J]
 * (serialFFT fft/ifft axis 0|1: pyfftw_fft.py:26-39,115-128) with fused zero-pad
 * (copy_to_padded slab.py:517-523), truncate + Nyquist fold (copy_from_padded slab.py:529-533,
 * :480-482), scaling (slab.py:320,483) and mask. */
typedef struct {
  int precision;
  int n;          /* transform length (logical, i.e. the padded length when padding) */
  long long B;    /* outer batch */
  int J;          /* inner contiguous extent */
  int inverse;    /* 0: exp(-i..), 1: exp(+i..); normalisation only through `scale` */
  int fold_mode;  /* store truncation: 0 none, 1 add mode -N/2 onto +N/2, 2 keep mode -N/2 only (line.py:189) */
  double scale;
  b200fft_side_t in, out;
  b200fft_mask_t mask;
  /* cross twiddle of a two-launch ("four-step") transform of a long axis, N = n1 * n2: the first launch (n = n1
   * over rows n2 apart, the n2 interleaved sub-columns side by side as J' = n2 * J columns) multiplies its output
   * k1 of sub-column x2 = column / cross_div by W_N^(x2 * k1) (conjugated for the inverse) with N = cross_n;
   * the second launch (n = n2, B = n1) needs nothing.  cross_n == 0: none. */
  int cross_n, cross_div;
} b200fft_strided_desc_t;

/* Batched real<->complex FFT along contiguous rows (serialFFT rfft/irfft axis -1:
 * pyfftw_fft.py:71-83,160-173).  Real rows: `rows` rows of n reals with pitch rpitch.  Complex
 * side: b = row index, chunked along k.  nk = complex entries stored (R2C: truncation
 * copy_from_padded axis 2, slab.py:535) or present (C2R: zero pad, slab.py:524-525); C2R ignores
 * the imaginary parts of k=0 and k=n/2 like FFTW / pocketfft. */
typedef struct {
  int precision;
  int n;          /* real length, even */
  long long rows;
  int nk;
  double scale;
  void* real_base;
  long long rpitch;
  b200fft_side_t cside;
  /* optional permutation of the rows on the complex side ("y-blocked" intermediate of single-rank slab plans):
   * with rm_block > 0 real row r = x * rm_period + y has its spectrum at complex row
   *     ((y / rm_block) * rm_planes + x) * rm_block + y % rm_block
   * i.e. the array [x][y][k] is kept as [y block][x][y in block][k].  A pass along x then finds its rows
   * rm_block * (row length) apart instead of rm_period * (row length), and a pass along y reads whole
   * contiguous blocks.  rm_block == 0: row r at complex row r. */
  long long rm_period, rm_block, rm_planes;
} b200fft_rows_desc_t;

/* Stream-ordered copy with cudaMemcpyDefault (either side may be host memory; pinned host memory
 * is copied by DMA without staging).  Used by the Python layer when callers pass numpy arrays,
 * which is how every reference caller passes data (host arrays, slab.py:214,349). */
B200FFT_API int b200fft_copy(void* dst, const void* src, size_t bytes, void* stream);
B200FFT_API int b200fft_stream_sync(void* stream);

B200FFT_API int b200fft_exec_strided(const b200fft_strided_desc_t* d, void* stream);
B200FFT_API int b200fft_exec_r2c(const b200fft_rows_desc_t* d, void* stream);
B200FFT_API int b200fft_exec_c2r(const b200fft_rows_desc_t* d, void* stream);
/* ---------------------------------------------------------------------------------------------
 * The caller of the transforms: pointwise operations of a pseudo-spectral
 * Navier-Stokes right-hand side, /root/reference/demo/spectral_dns_solver.py:53-77,87-98.  Vector fields
 * are [3][n] arrays (component-major, like the demo's U_hat / U / dU); wavenumbers come from three 1D device
 * vectors through the point's linear index, so no wavenumber mesh is read from HBM.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int precision;
  long long n0, n1, n2;       /* local complex shape (complex_shape()), C order */
  const void *kx, *ky, *kz;   /* device vectors of n0, n1, n2 reals: this rank's (scaled) wavenumbers
                                 (get_local_wavenumbermesh(scaled=True), slab.py:160-189) */
} b200fft_ns_mesh_t;
/* curl_hat = i K x u_hat  (demo :60-64) */
B200FFT_API int b200fft_ns_curl(const b200fft_ns_mesh_t* m, const void* u_hat, void* curl_hat, void* stream);
/* out = a x b on [3][npoints] real arrays (demo :53-58, before the forward transforms) */
B200FFT_API int b200fft_ns_cross(int precision, long long npoints, const void* a, const void* b, void* out, void* stream);
/* du (the transformed cross product) -> right-hand side: pressure projection and viscous term (demo :72-76);
 * with u_hat0 / u_hat1 given also the Runge-Kutta bookkeeping of the stage (demo :91-97) in the same pass:
 * u_hat1 += a_dt * rhs;  u_hat = last ? u_hat1 : u_hat0 + b_dt * rhs.  With u_hat0 == u_hat1 == NULL the
 * right-hand side is written back to du. */
B200FFT_API int b200fft_ns_rhs(const b200fft_ns_mesh_t* m, double nu, void* du, void* u_hat, const void* u_hat0, void* u_hat1,
                               double a_dt, double b_dt, int last, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Communicators: replace the mpi4py communicator (comm.Alltoall / Alltoallw / Sendrecv_replace /
 * Scatter / Send / Recv call sites listed in SURVEY.md section 2 row 7) by NCCL over NVLink.
 * The caller distributes the 128-byte unique id (host memory) out of band (torch.distributed,
 * MPI, a file ...) -- the role MPI_Init / comm.Split (pencil.py:192-193) play upstream.
 * ------------------------------------------------------------------------------------------- */
typedef struct b200fft_comm* b200fft_comm_t;
B200FFT_API int b200fft_comm_unique_id(void* id128);
B200FFT_API int b200fft_comm_create(b200fft_comm_t* comm, int nranks, int rank, const void* id128);
B200FFT_API int b200fft_comm_destroy(b200fft_comm_t comm);

/* ---------------------------------------------------------------------------------------------
 * Distributed transform plans: slab.R2C (slab.py:49-536), pencil.R2CX / R2CY
 * (pencil.py:145-1477) and line.R2C (line.py:41-340).  One plan per rank (SPMD); all ranks call
 * exec collectively in the same order.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int kind;        /* B200FFT_SLAB / PENCIL_X / PENCIL_Y / LINE */
  int precision;
  long long N[3];  /* global real mesh (line: N[0], N[1]; N[2] ignored) */
  int nranks;      /* comm.Get_size() */
  int rank;        /* comm.Get_rank() */
  int P1, P2;      /* pencil process grid (pencil.py:184-195); ignored otherwise */
  double padsize;  /* 3/2-rule pad factor; only 1.5 has kernels */
  int drop_nyquist;/* pencil communication='AlltoallN' layout (pencil.py:197-199,908-910) */
  int transport;   /* B200FFT_TRANSPORT_* */
  b200fft_comm_t comm;   /* slab / line: all ranks */
  b200fft_comm_t comm0;  /* pencil: ranks with equal rank / P1 */
  b200fft_comm_t comm1;  /* pencil: ranks with equal rank % P1 */
  int chunks;      /* pipeline depth of the exchange (the MPI collectives of slab.py:281-332,406-471
                      cut into `chunks` pieces, each overlapped with the FFT passes of the next
                      piece on a second stream); 0 = automatic, 1 = no overlap */
  int pipeline;    /* how a slab exchange is cut into pieces: B200FFT_PIPELINE_X by local x planes -- z(c), y(c) |
                      exchange(c), then one x pass; B200FFT_PIPELINE_KZ by kz ranges -- one z pass, then y(c) |
                      exchange(c) | x(c), so the exchange overlaps FFT passes on BOTH sides (three-stage pipeline;
                      receive layout is chunk-major); B200FFT_PIPELINE_AUTO (0, default): KZ for plain / 2/3-rule
                      R2C transforms over the copy engines whose per-peer message is >= 96 MB (measured faster at
                      2, 4 and 8 GPUs), X otherwise (3/2-rule, small meshes, NCCL, fused stores, C2C) */
  int layout;      /* single-rank slab.R2C plans: B200FFT_LAYOUT_YBLOCK (0, default) keeps the array between the
                      passes y-blocked and runs z, x, y (inverse y, x, z) so that no pass has rows megabytes apart;
                      B200FFT_LAYOUT_NATURAL runs z, y, x on [x][y][kz] like slab.py:366-370 (A/B measurements).
                      line.R2C plans: NATURAL also keeps columns of >= 128 KB in one launch instead of the two-launch
                      (four-step) form */
} b200fft_plan_desc_t;

typedef struct b200fft_plan* b200fft_plan_t;

B200FFT_API int b200fft_plan_create(b200fft_plan_t* plan, const b200fft_plan_desc_t* d);
B200FFT_API int b200fft_plan_destroy(b200fft_plan_t plan);
/* bytes of device scratch the plan owns (the reference's work_arrays, mpibase.py:61-131) */
B200FFT_API size_t b200fft_plan_workspace_bytes(b200fft_plan_t plan);
/* fftn / fft2 (slab.py:349-485, pencil.py:634-883,1228-1477, line.py:179-260):
 * u real_shape() [dealias 3/2: real_shape_padded()] -> fu complex_shape().  u is not modified. */
B200FFT_API int b200fft_exec_forward(b200fft_plan_t plan, const void* u, void* fu, int dealias, void* stream);
/* ifftn / ifft2 (slab.py:214-346, pencil.py:386-632,1001-1226, line.py:262-340).  fu is not modified. */
B200FFT_API int b200fft_exec_inverse(b200fft_plan_t plan, const void* fu, void* u, int dealias, void* stream);
/* Copy-engine transport (transport = B200FFT_TRANSPORT_P2P, slab plans): the all-to-all becomes
 * cudaMemcpyAsync pushes into the peers' receive buffers over NVLink (DMA engines, no SMs), ordered
 * by 32-bit sequence flags and stream memory operations, so that it overlaps the FFT passes of the
 * next chunk without competing for SMs.  After plan_create every rank calls _p2p_handles (fills
 * 256 bytes: CUDA IPC handles of its three work buffers and its flag words), the host exchanges
 * them (allgather, rank order) and every rank calls _p2p_connect with the nranks*256 bytes.
 * Replaces the same collectives as the NCCL path (slab.py:281-332,406-471).
 *
 * Fused transport (transport = B200FFT_TRANSPORT_STORE, slab plans; same handshake): no copy step at
 * all -- the FFT pass that produces the exchanged data (forward: the y pass, inverse: the x pass)
 * stores every peer's block straight into that peer's receive buffer through the IPC mapping
 * (st.global over NVLink / NVSwitch), tile by tile as the butterflies finish, so the transfer
 * overlaps the math inside ONE kernel; an exchange step only publishes / awaits the sequence flags.
 * The send buffer, its HBM write + read and the per-copy launch cost of the other transports vanish. */
B200FFT_API int b200fft_plan_p2p_handles(b200fft_plan_t plan, void* handles256);
B200FFT_API int b200fft_plan_p2p_connect(b200fft_plan_t plan, const void* all_handles);
/* number of kernels / NCCL groups the last exec launched (for bench.py's gpu_launches) */
B200FFT_API int b200fft_plan_last_launches(b200fft_plan_t plan, int* kernels, int* exchanges);
/* device time of the exchange phases of the last exec, if timing was enabled (ms; <0 if not) */
B200FFT_API int b200fft_plan_set_timing(b200fft_plan_t plan, int on);
B200FFT_API int b200fft_plan_last_phase_ms(b200fft_plan_t plan, float* fft_ms, float* exchange_ms);
/* per-step record of the last exec (timing on): type 0 strided C2C, 1 R2C, 2 C2R, 3 exchange; device
 * time in ms; algorithmic bytes = operand read once + result written once (SURVEY.md section 8d;
 * for an exchange: bytes sent to other ranks); transform length n; pass = index of the logical pass
 * the step belongs to (the chunks of a pipelined pass share it).  Returns the step count in *n. */
B200FFT_API int b200fft_plan_last_steps(b200fft_plan_t plan, int max, int* n, int* type, float* ms, double* bytes, int* len,
                                        int* pass);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_H */
