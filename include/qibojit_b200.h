/*
 * qibojit_b200.h -- C ABI of the B200-native state-vector gate path.
 *
 * One entry point per kernel that qibojit binds for this path.  "Replaces" lines cite the
 * reference interface (numba kernel in /root/reference/src/qibojit/custom_operators/ and the
 * CuPy RawKernel + launcher in .../custom_operators/raw_kernels.py, .../backends/gpu.py).
 *
 * Conventions
 *  - `state` is a DEVICE pointer to 2^nqubits amplitudes, interleaved (re, im);
 *    dtype QJ_C64 = complex64 (2 x float), QJ_C128 = complex128 (2 x double).
 *    It must be 32-byte aligned (any cudaMalloc / torch allocation is).
 *  - `gate` is a HOST pointer to the row-major gate buffer in the state dtype (the
 *    reference hands a device array, gpu.py:936; here the library stages it itself so that
 *    no gate costs a device allocation or a synchronisation).
 *  - `qubits` is a HOST int32 array: sorted ascending index-bit positions (n-1-q) of
 *    controls U targets, exactly cpu.py:565-569.  NULL / nactive == ntargets selects the
 *    kernel without controls (cpu.py:612-616, 631-635).
 *  - every call is asynchronous on the handle's stream; qj_sync() waits.  All kernels work
 *    in place; nothing state-sized is ever allocated by the library.
 *  - return value: 0 = ok, negative = error (QJ_ERR_*); qj_last_error() gives the message
 *    for the calling thread.  Handles are per device and may be used from one thread at a
 *    time (the reference drives one device per joblib thread, gpu.py:688-694).
 */
#ifndef QIBOJIT_B200_H
#define QIBOJIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QJ_C64 0
#define QJ_C128 1

#define QJ_OK 0
#define QJ_ERR_INVALID (-1) /* bad argument (ValueError upstream, gpu.py:989-993)   */
#define QJ_ERR_CUDA (-2)    /* CUDA runtime error (RuntimeError upstream, gpu.py:733) */
#define QJ_ERR_NODEVICE (-3)
#define QJ_ERR_UNSUPPORTED (-4)

#define QJ_MAX_QUBITS 48       /* index bits addressable by the kernels            */
#define QJ_MAX_TARGETS 10      /* dense k-target gate limit (reference GPU: 7, gpu.py:48) */

typedef struct qj_handle qj_handle;

/* ---- handle / runtime ------------------------------------------------------------ */
/* stream: the cudaStream_t every call of this handle is enqueued on (e.g. torch's current
 * stream); NULL is the CUDA default stream. */
int qj_create(int device, void *stream, qj_handle **out);
int qj_destroy(qj_handle *h);
int qj_set_stream(qj_handle *h, void *stream);
int qj_sync(qj_handle *h);
const char *qj_last_error(void);
const char *qj_version(void);
/* number of kernels launched through this handle since creation (bench: gpu_launches) */
int64_t qj_launch_count(qj_handle *h);
/* kernel routing: 0 = automatic, 1 = force the register ("direct") kernels,
 * 2 = force the shared-memory tile kernel where it applies.  For tests and profiling. */
int qj_set_route(qj_handle *h, int route);

/* ---- state preparation -------------------------------------------------------------
 * replaces ops.initial_state_vector (ops.py:14-18), initial_state_kernel
 * (raw_kernels.py:548-558) + cp.zeros (gpu.py:584-600).                                 */
int qj_initial_state(qj_handle *h, void *state, int dtype, int nqubits);

/* ---- one-target kernels ------------------------------------------------------------
 * replace {,multicontrol_}apply_gate / apply_x / apply_y / apply_z / apply_z_pow
 * (gates.py:16-114, raw_kernels.py:128-221 and 293-391; launcher gpu.py:1009-1036).
 * m = index bit of the target.  gate: 2x2 (apply_gate) or one scalar (apply_z_pow).      */
int qj_apply_gate(qj_handle *h, void *state, int dtype, int nqubits, int m, const void *gate,
                  const int32_t *qubits, int nactive);
int qj_apply_x(qj_handle *h, void *state, int dtype, int nqubits, int m,
               const int32_t *qubits, int nactive);
int qj_apply_y(qj_handle *h, void *state, int dtype, int nqubits, int m,
               const int32_t *qubits, int nactive);
int qj_apply_z(qj_handle *h, void *state, int dtype, int nqubits, int m,
               const int32_t *qubits, int nactive);
int qj_apply_z_pow(qj_handle *h, void *state, int dtype, int nqubits, int m, const void *phase,
                   const int32_t *qubits, int nactive);

/* multiply every amplitude by one complex scalar (HOST pointer, state dtype): the form a
 * diagonal gate takes on a shard whose rank bits fix all of its qubits (distributed rule iii;
 * no single reference kernel -- the reference rebuilds the full vector on the host for such
 * gates, gpu.py:1478-1495). */
int qj_apply_phase(qj_handle *h, void *state, int dtype, int nqubits, const void *phase);

/* ---- two-target kernels --------------------------------------------------------------
 * replace {,multicontrol_}apply_two_qubit_gate / apply_swap / apply_fsim
 * (gates.py:118-254, raw_kernels.py:224-290 and 394-477; launcher gpu.py:1038-1076).
 * m1 < m2 are the target index bits, swap_targets as computed in cpu.py:622-629.
 * gate: 4x4 (two_qubit_gate) or the 5-vector of matrices.py:62-70 (fsim).                */
int qj_apply_two_qubit_gate(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2,
                            int swap_targets, const void *gate, const int32_t *qubits,
                            int nactive);
int qj_apply_swap(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2,
                  const int32_t *qubits, int nactive);
int qj_apply_fsim(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2,
                  int swap_targets, const void *gate, const int32_t *qubits, int nactive);

/* ---- k-target dense kernel (k >= 1; fused blocks use the same entry) ---------------------
 * replaces apply_three/four/five_qubit_gate_kernel and apply_multi_qubit_gate_kernel
 * (gates.py:266-424, raw_kernels.py:480-518; launcher gpu.py:975-1007).
 * targets: HOST int64 single-bit masks in REVERSED target order (cpu.py:594-596):
 * targets[u] is the index bit that matrix-index bit u addresses.                          */
int qj_apply_multi_qubit_gate(qj_handle *h, void *state, int dtype, int nqubits,
                              const void *gate, const int32_t *qubits, int nactive,
                              const int64_t *targets, int ntargets);

/* ---- measurement -----------------------------------------------------------------------
 * qj_collapse_state replaces ops.collapse_state / collapse_state_normalized (ops.py:47-79),
 * collapse_state_kernel (raw_kernels.py:521-545) + the cupy normalisation (gpu.py:640-642).
 * qubits: HOST int32 bit positions [n-q-1 for q in reversed(sorted measured)] (cpu.py:550).  */
int qj_collapse_state(qj_handle *h, void *state, int dtype, int nqubits, const int32_t *qubits,
                      int ntargets, int64_t result, int normalize);
/* squared 2-norm of the state into a host double (synchronises). */
int qj_norm2(qj_handle *h, const void *state, int dtype, int nqubits, double *out);
/* max_i |state[i] - (ref_re + i ref_im)| into a host double (synchronises): the distance from a
 * constant vector -- the closed form of the reference's own benchmark circuit QFT|0..0> =
 * 2^(-n/2) everywhere (benchmarks/abstract.py:86-104) -- at sizes no host copy can check.   */
int qj_max_deviation(qj_handle *h, const void *state, int dtype, int nqubits, double ref_re,
                     double ref_im, double *out);
/* replaces qibo Backend.calculate_probabilities (call sites gpu.py:751-768): probs is a
 * DEVICE array of 2^nmeas reals (float for C64, double for C128); bits[j] = index bit of the
 * j-th measured qubit, output index has j = 0 as its most significant bit.               */
int qj_calculate_probabilities(qj_handle *h, const void *state, int dtype, int nqubits,
                               const int32_t *bits, int nmeas, void *probs);
/* replaces ops.measure_frequencies (ops.py:86-108; called by cpu.py:383-394 and, on the
 * host, by gpu.py:26,61).  frequencies: DEVICE int64[2^nqubits] (accumulated into);
 * probs: DEVICE real[2^nqubits] (real_dtype QJ_C64 -> float, QJ_C128 -> double).
 * Same MT19937 streams as numba: nthreads independent chains.                             */
int qj_measure_frequencies(qj_handle *h, int64_t *frequencies, const void *probs,
                           int real_dtype, int64_t nshots, int nqubits, int64_t seed,
                           int nthreads);
/* replaces qibo Backend.sample_shots (gpu.py:770-773): inverse-CDF sampling of `nshots`
 * indices from probs (DEVICE real[2^nqubits]) with HOST uniforms u[nshots] in [0,1) drawn by
 * the caller's generator; shots: DEVICE int64[nshots].  cdf_scratch: DEVICE double[2^nqubits]. */
int qj_sample_shots(qj_handle *h, const void *probs, int real_dtype, int nqubits,
                    const double *uniforms, int64_t nshots, int64_t *shots, double *cdf_scratch);

/* ---- distributed state: global <-> local qubit swap ----------------------------------------
 * replaces ops.swap_pieces (ops.py:131-137; driver loop gpu.py:1497-1507).  `local` is this
 * rank's shard of 2^nlocal amplitudes, `peer` a pointer to the partner rank's shard that is
 * addressable from this device (NVLink peer mapping / symmetric memory).  The rank whose
 * global bit is 0 passes is_upper = 0.  Each rank moves half of the exchanged amplitudes:
 * together the two calls perform piece0[i + 2^m] <-> piece1[i] for all i with bit m clear. */
int qj_swap_pieces_peer(qj_handle *h, void *local, void *peer, int dtype, int nlocal, int m,
                        int is_upper);
/* The same exchange for SEVERAL qubits at once over peer memory (the all-to-all of a multi-qubit
 * global<->local swap, one call per peer): this rank's sub-block whose index bits `bits`
 * (ascending) spell `local_value` trades places with the peer's sub-block that spells
 * `peer_value`.  The two ranks of a pair each call it with their own `part` of `nparts` (0 / 1 of
 * 2): together they move every amplitude once.  No staging buffer, no pack / unpack pass.      */
int qj_swap_bits_peer(qj_handle *h, void *local, void *peer, int dtype, int nlocal, const int32_t *bits,
                      int nbits, int local_value, int peer_value, int part, int nparts);
/* CUDA IPC plumbing of the peer path (one process per GPU): export the allocation holding `ptr`
 * (64-byte handle + offset of `ptr` inside it), map a peer's allocation, unmap it.             */
int qj_ipc_export(const void *ptr, void *handle64, int64_t *offset);
int qj_ipc_open(const void *handle64, void **base_out);
int qj_ipc_close(void *base);
/* Stream-ordered handshake with `npeers` peers over mapped 32-bit flag words (the device-side
 * replacement of the host barrier between the reference's per-piece kernels and its piece swaps,
 * gpu.py:1497-1507, for one process per GPU): in stream order, publish epochs[i] into
 * remote_slots[i] (the peer's flag slot for this rank), then wait until local_flags[src_ranks[i]]
 * reaches epochs[i].  Work enqueued before the call is complete and visible to a peer that has
 * seen the epoch.  A peer that does not answer within `timeout_seconds` (<= 0: 30 s) fails the
 * launch (surfaces as a CUDA error at the next synchronisation) instead of hanging the device.  */
int qj_peer_handshake(qj_handle *h, void *local_flags, void *const *remote_slots, const int32_t *src_ranks,
                      const uint32_t *epochs, int npeers, double timeout_seconds);
/* Asynchronous copy on the handle's stream by the copy engines (no SM is used, so it runs under a
 * pass kernel at full rate); either side may be a peer allocation mapped with qj_ipc_open.      */
int qj_copy_async(qj_handle *h, void *dst, const void *src, int64_t bytes);
/* Staged variant for transports without peer mapping (NCCL send/recv of chunks):
 * pack: gather the half of `local` that leaves (bit m == 1 - is_upper) for amplitudes
 * [chunk_begin, chunk_begin + chunk_len) of the half-shard into contiguous `buf`;
 * unpack: scatter a received chunk into the same slots.                                    */
int qj_swap_pack(qj_handle *h, const void *local, void *buf, int dtype, int nlocal, int m,
                 int is_upper, int64_t chunk_begin, int64_t chunk_len);
int qj_swap_unpack(qj_handle *h, void *local, const void *buf, int dtype, int nlocal, int m,
                   int is_upper, int64_t chunk_begin, int64_t chunk_len);
/* Several global qubits at once (one all-to-all among the 2^nbits ranks that differ in the
 * exchanged rank bits instead of nbits pairwise half-shard swaps: (2^nbits - 1) / 2^nbits of a
 * shard crosses the links instead of nbits / 2).  pack: gather the sub-block of `local` whose
 * index bits bits[i] (ascending) equal bit i of `value` -- the amplitudes that go to the rank
 * whose exchanged rank bits spell `value` -- elements [chunk_begin, chunk_begin + chunk_len)
 * of that sub-block, into contiguous `buf`; unpack: scatter what that rank sent into the same
 * slots.  Generalises ops.swap_pieces (ops.py:131-137) to the piece transposition of
 * ops.transpose_state / MultiGpuOps.to_pieces (gpu.py:1437-1465).                            */
#define QJ_MAX_GLOBAL_SWAP 6
int qj_swap_pack_bits(qj_handle *h, const void *local, void *buf, int dtype, int nlocal,
                      const int32_t *bits, int nbits, int value, int64_t chunk_begin, int64_t chunk_len);
int qj_swap_unpack_bits(qj_handle *h, void *local, const void *buf, int dtype, int nlocal,
                        const int32_t *bits, int nbits, int value, int64_t chunk_begin, int64_t chunk_len);

/* ---- multi-gate passes ("tile programs") ------------------------------------------------------
 * replaces a whole gate queue per device: the loop `for gate in queue: apply_gate(...)` of
 * qibo's Backend.execute_circuit and of MultiGpuOps.apply_gates (gpu.py:1467-1476), where every
 * gate is one pass over the state.  A pass stages tiles of 2^nlocal amplitudes (the index bits
 * `local_bits`, which must start with the contiguous run 0,1,..,r-1) in shared memory and applies
 * all of its ops to the tile before writing it back: ONE HBM pass for the whole op list.
 * Inside a pass the ops are grouped into ROUNDS: in a round every thread holds the 2^nreg
 * amplitudes spanned by the round's register bits (local bits; complex128: nreg = 4,
 * complex64: nreg = 5 and index bit 0 must be one of them) and applies the round's ops to them.
 *   QJ_OPK_DENSE1 / QJ_OPK_DENSE2: the arithmetic of {,multicontrol_}apply_gate_kernel and
 *     {,multicontrol_}apply_two_qubit_gate_kernel (gates.py:16-38, 118-193); targets must be
 *     register bits of the round, targets[j] is the index bit addressed by matrix-index bit j
 *     (the reversed-target convention of cpu.py:594-596); controls may be any index bits.
 *   QJ_OPK_DIAG: multiply amplitude i by table[sum_j bit(i, targets[j]) << j] when every control
 *     bit of i is 1 -- a product of diagonal gates (apply_z / apply_z_pow, gates.py:82-114, and
 *     any other diagonal matrix) merged on the host; its bits may be any index bits.
 * `data` (HOST, complex in the state dtype) holds the row-major matrices and the tables; every op
 * owns its own range starting at data_offset.  Rounds of a pass and ops of a round are applied
 * in order.                                                                                    */
#define QJ_OPK_DENSE1 1
#define QJ_OPK_DENSE2 2
#define QJ_OPK_DIAG 3
#define QJ_MAX_DIAG_BITS 12
#define QJ_MAX_LOCAL_BITS 16
#define QJ_MAX_REG_BITS 8

typedef struct qj_op_desc {
    int32_t kind;
    int32_t ntargets;                 /* dense: 1 or 2; diag: number of table-index bits (0..12) */
    int32_t ncontrols;
    int32_t reserved;
    int64_t data_offset;              /* in complex elements */
    int32_t targets[QJ_MAX_DIAG_BITS];
    int32_t controls[QJ_MAX_QUBITS];
} qj_op_desc;

typedef struct qj_round_desc {
    int32_t nreg;                     /* 4 (complex128) / 5 (complex64) */
    int32_t reserved;
    int64_t first_op;                 /* range of this round in the `ops` array */
    int64_t nops;
    int32_t reg_bits[QJ_MAX_REG_BITS]; /* strictly ascending index bits, a subset of local_bits */
} qj_round_desc;

typedef struct qj_pass_desc {
    int32_t nlocal;                   /* 6 .. 12 (complex128) / 13 (complex64) */
    int32_t reserved;
    int64_t first_round;              /* range of this pass in the `rounds` array */
    int64_t nrounds;
    int32_t local_bits[QJ_MAX_LOCAL_BITS]; /* strictly ascending index bits */
} qj_pass_desc;

typedef struct qj_program qj_program;

int qj_program_create(qj_handle *h, int dtype, int nqubits, const qj_pass_desc *passes, int npasses,
                      const qj_round_desc *rounds, int64_t nrounds, const qj_op_desc *ops,
                      int64_t nops, const void *data, int64_t ndata, qj_program **out);
/* apply every pass of the program to `state` (DEVICE pointer, 2^nqubits amplitudes) */
int qj_program_run(qj_handle *h, const qj_program *p, void *state);
/* one kernel launch of the program (bench/profiling: per-pass timing) */
int qj_program_run_launch(qj_handle *h, const qj_program *p, void *state, int launch);
/* Launches [first_launch, first_launch + nlaunches) (nlaunches < 0: to the end).  QJ_RUN_ZERO_INPUT:
 * the state is |0...0> on entry and need not hold it -- the first of these launches writes every
 * amplitude without reading any (`initial_state_vector`, ops.py:14-18, fused into the first pass:
 * one state-sized write and one read less per circuit).                                          */
#define QJ_RUN_ZERO_INPUT 1
int qj_program_run_ex(qj_handle *h, const qj_program *p, void *state, int first_launch, int nlaunches,
                      int flags);
/* One launch restricted to the tiles [tile_begin, tile_begin + tile_count): a caller that knows a
 * sub-block of the state is complete (the distributed layer pipelines the last pass before a qubit
 * exchange against the exchange itself, sub-block by sub-block) runs the pass on that sub-block
 * alone.  Tiles are numbered by the index bits outside the tile, least significant first;
 * `qj_program_launch_geometry` fills out[0..11] = {tile bits, run bits, high tile bits, number of
 * tiles, the high tile bits' index bits (8 entries, -1 = unused)}.                             */
int qj_program_run_tiles(qj_handle *h, const qj_program *p, void *state, int launch, int64_t tile_begin,
                         int64_t tile_count);
/* The same launch reading the tiles from `state` and storing them at the same positions of `out`
 * (another buffer of the same size): the pass that precedes a qubit exchange deposits the sub-block
 * this rank keeps straight into the buffer the exchange fills.                                  */
int qj_program_run_tiles_to(qj_handle *h, const qj_program *p, const void *state, void *out, int launch,
                            int64_t tile_begin, int64_t tile_count);
int qj_program_launch_geometry(const qj_program *p, int launch, int64_t *out);
int qj_program_stats(const qj_program *p, int64_t *nlaunches, int64_t *nrounds, int64_t *nmops);
int qj_program_destroy(qj_handle *h, qj_program *p);

/* Inspection (CPU tests of the host-side encoder; no device needed): the encoded image of the
 * same arguments -- program units (16 bytes each), phase tables / per-thread factors, and per
 * launch QJ_LAUNCH_INFO_FIELDS int64 values {tile bits, run bits, high bits, tiles, first unit,
 * units, per-tile factors, shared-memory bytes, high bit positions[8]}.
 * tests/pass_emulator.py interprets it.                                                       */
#define QJ_LAUNCH_INFO_FIELDS 16
typedef struct qj_program_image qj_program_image;
int qj_program_encode(int dtype, int nqubits, const qj_pass_desc *passes, int npasses,
                      const qj_round_desc *rounds, int64_t nrounds, const qj_op_desc *ops,
                      int64_t nops, const void *data, int64_t ndata, qj_program_image **out);
int qj_program_image_sizes(const qj_program_image *img, int64_t *nlaunches, int64_t *blob_bytes,
                           int64_t *table_bytes);
int qj_program_image_read(const qj_program_image *img, void *blob_out, void *tables_out,
                          int64_t *launch_info);
int qj_program_image_destroy(qj_program_image *img);

#ifdef __cplusplus
}
#endif
#endif /* QIBOJIT_B200_H */
