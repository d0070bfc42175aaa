/* pffrg.h -- C ABI of libpffrg: the B200-native pf-FRG flow-equation core.
 *
 * This is the drop-in boundary for SpinParser's FrgCore plugin seam. The reference host code (task/lattice/model
 * parsing, symmetry reduction, measurements, HDF5 output) stays as it is and drives a flow core through
 *     FrgCore *FrgCoreFactory::newFrgCore(identifier, spinModel, measurements, options)   src/FrgCoreFactory.hpp:40
 *     virtual void FrgCore::computeStep()                                                  src/FrgCore.hpp:77
 *     virtual void FrgCore::finalizeStep(float newCutoff)                                  src/FrgCore.hpp:86
 *     EffectiveAction::{cutoff, isDiverged()} and the vertex arrays the measurements read  src/EffectiveAction.hpp:40-59
 * Each entry point below names the reference interface it replaces. INTEGRATION.md shows the adapter class
 * (`B200FrgCore : FrgCore`) a SpinParser maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C types only; the caller owns every host buffer it passes; the library owns all device memory.
 *  - every function returns PFFRG_OK (0) or a negative error code; pffrg_last_error() describes the last failure.
 *  - a handle is driven by one host thread; one handle per GPU (one process per GPU for multi-GPU runs).
 *  - host arrays use the REFERENCE memory layout and either float (the reference's type) or double:
 *        v2   [Nw]                                     src/SU2/SU2VertexSingleParticle.hpp
 *        v4_c [su][t][rid], su = so(so+1)/2+uo, so>=uo src/SU2/SU2VertexTwoParticle.hpp:595-605 (SU2: c = S,D; XYZ: X,Y,Z,D)
 *        v4   [su][t][mu][nu][rid]                     src/TRI/TRIVertexTwoParticle.hpp:65-71    (TRI: one array)
 *    All arithmetic on the device is FP64.
 *  - there is no CPU fallback: every compute entry point fails with PFFRG_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PFFRG_H
#define PFFRG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFFRG_ABI_VERSION 3

typedef enum pffrg_status
{
	PFFRG_OK = 0,
	PFFRG_ERR_ARGUMENT = -1,   /* invalid descriptor / null pointer / wrong size */
	PFFRG_ERR_CUDA = -2,       /* CUDA runtime or launch failure, or no usable GPU */
	PFFRG_ERR_NCCL = -3,       /* NCCL failure */
	PFFRG_ERR_STATE = -4,      /* call sequence violated (e.g. compute before set_state) */
	PFFRG_ERR_UNSUPPORTED = -5 /* problem exceeds a compiled-in limit */
} pffrg_status;

/* identifier strings of FrgCoreFactory::newFrgCore, src/FrgCoreFactory.cpp:25-51 */
typedef enum pffrg_core { PFFRG_CORE_SU2 = 0, PFFRG_CORE_XYZ = 1, PFFRG_CORE_TRI = 2 } pffrg_core;
typedef enum pffrg_dtype { PFFRG_F32 = 0, PFFRG_F64 = 1 } pffrg_dtype;

/* Problem description: everything the hot path reads from FrgCommon::{frequency(),lattice()} (src/FrgCommon.hpp:26-49)
 * and the core options of src/SU2/SU2FrgCore.cpp:20-29. All tables are copied during pffrg_create. */
typedef struct pffrg_desc
{
	int32_t abi_version;          /* PFFRG_ABI_VERSION */
	int32_t core;                 /* pffrg_core */
	int32_t n_frequencies;        /* Nw: FrequencyDiscretization::size */
	const double *frequencies;    /* [Nw] strictly ascending positive mesh points (FrequencyDiscretization::_data) */
	int32_t n_sites;              /* L: Lattice::size (number of representative sites) */
	const int32_t *sites_rid;     /* [L]   Lattice::getSites()[j].rid                  src/Lattice.hpp:491 */
	const int32_t *sites_perm;    /* [L*3] Lattice::getSites()[j].spinPermutation[k]  (0=X 1=Y 2=Z) */
	const int32_t *inverted_rid;  /* [L]   Lattice::getInvertedSites()[j].rid          src/Lattice.hpp:481 */
	const int32_t *inverted_perm; /* [L*3] */
	const int32_t *overlap_offsets; /* [L+1] prefix sums of Lattice::getOverlap(rid).size  src/Lattice.hpp:46-150,470 */
	const int32_t *overlap_rid1;  /* [overlap_offsets[L]] LatticeOverlap::rid1 */
	const int32_t *overlap_rid2;  /* [overlap_offsets[L]] LatticeOverlap::rid2 */
	const int32_t *overlap_perm1; /* [overlap_offsets[L]*3] transformed{X,Y,Z}1 */
	const int32_t *overlap_perm2; /* [overlap_offsets[L]*3] transformed{X,Y,Z}2 */
	int32_t n_range;              /* number of sites j in Lattice::getRange(0)         src/Lattice.hpp:503-523 */
	const int32_t *range_fwd_rid; /* [n_range] Lattice::symmetryTransform(zero, j)     src/Lattice.hpp:397-402 */
	const int32_t *range_inv_rid; /* [n_range] Lattice::symmetryTransform(j, zero) */
	double spin_length;           /* SU2 option "spin" (ignored by XYZ/TRI)            src/SU2/SU2FrgCore.cpp:20,25 */
	int32_t device;               /* CUDA device ordinal for this handle */
} pffrg_desc;

typedef struct pffrg_context *pffrg_handle;

/* per-step statistics (replaces the LoadManager's per-rank timing gather, src/lib/LoadManager.hpp:636-665) */
typedef struct pffrg_stats
{
	double ms_v2_flow;      /* device time of the self-energy flow kernel */
	double ms_node_table;   /* device time of the quadrature node table kernel */
	double ms_v4_flow;      /* device time of the vertex flow kernel (the hot kernel) */
	double ms_finalize;     /* device time of the last finalize (Euler update + exchange) */
	double ms_exchange;     /* device time of the NCCL exchange inside the last finalize */
	int64_t kernel_evals;   /* quadrature nodes (kernel evaluations) this rank processed in the last step, all 3 channels */
	int64_t kernel_evals_t; /* of which t-channel evaluations */
	int64_t items;          /* work items (frequency triples) this rank processed */
	double alg_bytes;       /* algorithmic gather+output bytes of this rank's share */
	double alg_flops;       /* algorithmic FP64 flops of this rank's share */
	int32_t launches;       /* kernels launched by the last compute_step + finalize_step */
	int32_t jit_rpa;        /* 1 when the lattice-specialised (run-time compiled) flow kernel is in use */
	double jit_compile_ms;  /* time spent generating + compiling it in pffrg_create */
	int32_t threads;        /* launch shape of the flow kernel: threads per CTA, */
	int32_t smem_bytes;     /*   dynamic shared memory per CTA, */
	int32_t node_batch;     /*   quadrature nodes per gather batch, */
	int32_t rpa_batch;      /*   t-channel nodes staged per RPA phase, */
	int32_t rpa_warps;      /*   warps sharing the RPA instruction stream(s), */
	int32_t min_blocks;     /*   CTAs per SM the kernel was compiled for */
	int32_t autotuned_shapes; /* launch shapes compiled and timed in pffrg_create (0/1: no autotuning) */
	int32_t sub_ctas;       /* work items per CTA of the run-time compiled kernel (sub-CTAs sharing one RPA phase); fills the former tail padding */
	int32_t gram_rows;      /* > 0: the RPA phase runs in its Gram form (rpaGram) with this many rows of the Gram matrix per block */
	int32_t rpa_terms_merged; /* overlap terms after merging equal (rid1, rid2[, permutations]) pairs: what the RPA phase walks */
	double exec_flops;      /* FP64 flops the kernels execute for this rank's share: as alg_flops, but the RPA term counted as implemented
	                           (merged terms per node for the straight-line / word-stream forms; L^2 per node + merged terms per RPA phase for the Gram form) */
	int32_t gather_threads; /* > 0: warp-specialised flow kernel (gather / RPA / producer warp groups in one CTA) with this many gather threads; */
	int32_t producer_warps; /*   warps that build the access-buffer tables ahead of the gather warps (0: the gather warps build their own) */
} pffrg_stats;

/* library / environment ------------------------------------------------------------------------------------------ */
int pffrg_abi_version(void);
/* description of the last error on this thread (never NULL) */
const char *pffrg_last_error(void);
/* number of usable CUDA devices (0 when there is no GPU; never fails) */
int pffrg_device_count(void);

/* lifetime: replaces FrgCoreFactory::newFrgCore + the {SU2,XYZ,TRI}FrgCore constructors/destructors
 * (src/FrgCoreFactory.cpp:25-51, src/SU2/SU2FrgCore.cpp:17-87) ------------------------------------------------------ */
int pffrg_create(const pffrg_desc *desc, pffrg_handle *out);
int pffrg_destroy(pffrg_handle h);

/* sizes of the host arrays in reference layout */
int pffrg_num_vertex_arrays(pffrg_handle h);     /* 2 (SU2) / 4 (XYZ) / 1 (TRI) */
int64_t pffrg_vertex_array_length(pffrg_handle h); /* elements per array: NF*L (SU2/XYZ), NF*16*L (TRI) */
int64_t pffrg_num_items(pffrg_handle h);         /* NF = Nw*Nw*(Nw+1)/2 work items */

/* multi-GPU: replaces HMP::LoadManager's MPI distribution and broadcasts (src/lib/LoadManager.hpp:581-649,240).
 * One process per GPU. Rank 0 calls pffrg_comm_unique_id and ships the id to all ranks by any host transport
 * (the reference would use MPI_Bcast); then every rank calls pffrg_comm_init. Without these calls the handle is a
 * single-GPU core. Work items are split into contiguous, cost-balanced ranges (pffrg_item_range). */
#define PFFRG_UNIQUE_ID_BYTES 128
int pffrg_comm_unique_id(void *id_out /* PFFRG_UNIQUE_ID_BYTES */);
int pffrg_comm_init(pffrg_handle h, const void *id, int rank, int n_ranks);
int pffrg_item_range(pffrg_handle h, int64_t *begin, int64_t *end); /* this rank's items in the current step */
/* The partition itself, as a pure host function (no GPU needed): bounds[r] .. bounds[r+1] are the work items of rank r at
 * `cutoff`; `rpa_terms` = number of distinct overlap terms (after merging duplicates; overlap_offsets[L] is a fine estimate).
 * Item costs follow the exact quadrature node counts of src/lib/Integrator.hpp:138-287, so every rank gets the same load. */
int pffrg_plan_partition(int core, int n_frequencies, const double *frequencies, int n_sites, int64_t rpa_terms, double cutoff,
                         int n_ranks, int64_t *bounds /* [n_ranks + 1] */);
/* The same with feedback, as the library applies it from the second step of a multi-GPU run on (replaces the throughput-adaptive
 * chunk sizes of src/lib/LoadManager.hpp:796-852): `prev_bounds` [n_ranks + 1] and `prev_ms` [n_ranks] are the boundaries and the
 * measured flow-kernel times of the previous step; the modelled cost of an item is weighted by the measured time per modelled
 * unit of the previous interval it lies in. PFFRG_BALANCE=0 keeps the static split. */
int pffrg_plan_partition_feedback(int core, int n_frequencies, const double *frequencies, int n_sites, int64_t rpa_terms, double cutoff,
                                  int n_ranks, const int64_t *prev_bounds, const double *prev_ms, int64_t *bounds /* [n_ranks + 1] */);

/* state transfer: replaces direct access to {SU2,XYZ,TRI}EffectiveAction's arrays and EffectiveAction::cutoff
 * (src/SU2/SU2EffectiveAction.hpp:38-60, src/EffectiveAction.hpp:59). `v4` holds pffrg_num_vertex_arrays pointers. */
int pffrg_set_state(pffrg_handle h, double cutoff, const void *v2, const void *const *v4, int dtype);
int pffrg_get_state(pffrg_handle h, double *cutoff, void *v2, void *const *v4, int dtype);
/* Multi-GPU runs in which every rank holds the same host state (as the MPI ranks of the reference do): each rank uploads only the
 * rows [begin, end) of pffrg_upload_slice (an even split of the work items) and the slices are distributed over NVLink (peer-memory
 * stores), instead of every rank pushing the whole vertex through PCIe. Collective; falls back to pffrg_set_state on one GPU.
 * pffrg_get_state_slice downloads the rows [begin, end) of the state into the corresponding rows of the host arrays (the other rows
 * are left untouched): a rank that only needs its share of the result (or the master rank, the whole) reads just that. */
int pffrg_upload_slice(pffrg_handle h, int64_t *begin, int64_t *end);
int pffrg_set_state_sharded(pffrg_handle h, double cutoff, const void *v2, const void *const *v4, int dtype);
int pffrg_get_state_slice(pffrg_handle h, double *cutoff, void *v2, void *const *v4, int dtype, int64_t begin, int64_t end);
/* Initial condition built on the device (SU2EffectiveAction(cutoff, spinModel, core), src/SU2/SU2EffectiveAction.hpp:38-60; XYZ
 * :37-62, TRI :38-63): every frequency entry of channel c at representative r is bare[c * n_sites + r] -- the bare coupling already
 * divided by the normalization (and multiplied by 1/4 for XYZ/TRI) -- and the self energy is zero. Replaces building and uploading
 * the full host arrays. */
int pffrg_set_initial_condition(pffrg_handle h, double cutoff, const double *bare /* [pffrg_num_channels * n_sites] */);
/* the flow of the last compute_step (FrgCore::flow(), src/FrgCore.hpp:103-106); gathers from all ranks when needed */
int pffrg_get_flow(pffrg_handle h, void *v2_flow, void *const *v4_flow, int dtype);

/* FrgCore::computeStep (src/SU2/SU2FrgCore.cpp:89-109): self-energy flow, then vertex flow of this rank's items.
 * *diverged (optional) is set to 1 when the flow contains NaN (EffectiveAction::isDiverged,
 * src/SU2/SU2EffectiveAction.hpp:212-230); divergence is not an error. */
int pffrg_compute_step(pffrg_handle h, int *diverged);
/* FrgCore::finalizeStep (src/SU2/SU2FrgCore.cpp:111-137): state += (new_cutoff - cutoff) * flow, cutoff = new_cutoff,
 * then the updated vertex is exchanged between ranks (ncclBroadcast group == the reference's MPI_Bcast). */
int pffrg_finalize_step(pffrg_handle h, double new_cutoff);
/* Static spin-spin correlations chi_c[rid] of the CURRENT state, computed on the device: replaces the work item of
 * {SU2,XYZ,TRI}MeasurementCorrelation::_calculateCorrelation (src/SU2/SU2MeasurementCorrelation.cpp:77-160; XYZ :98-205, TRI
 * :147-296), which the reference runs single-threaded once per cutoff step. `chi` receives channels x representatives doubles,
 * channel-major ("ValueSuperbundle" order: SU2 {spin, density}, XYZ {x, y, z, density}, TRI 4*mu+nu); the caller maps
 * representatives to lattice sites with Lattice::symmetryTransform exactly as src/SU2/SU2MeasurementCorrelation.cpp:161-176 does. */
int pffrg_num_channels(pffrg_handle h);          /* 2 (SU2) / 4 (XYZ) / 16 (TRI) */
int pffrg_measure_correlation(pffrg_handle h, double *chi /* [pffrg_num_channels * n_sites] */);
/* block until all device work of this handle has finished */
int pffrg_synchronize(pffrg_handle h);

/* restrict the next compute_step to an explicit item range (tests, bounded benchmarks); end<=begin restores the default */
int pffrg_set_item_range(pffrg_handle h, int64_t begin, int64_t end);

int pffrg_get_stats(pffrg_handle h, pffrg_stats *out);

/* raw CUDA stream (cudaStream_t) the handle launches on, for callers that time with their own events */
void *pffrg_stream(pffrg_handle h);

/* Generate and compile (NVRTC, sm_100a) the lattice-specialised flow kernel for a problem without touching a GPU:
 * a build-time check that the run-time compiled path works for this lattice. *cubin_bytes receives the code size. */
int pffrg_jit_compile_check(const pffrg_desc *desc, int64_t *cubin_bytes);

/* Bilinear term tables of the TRI core as the kernels use them, derived from the spin algebra (not stored): replaces the
 * machine-generated statement lists of src/TRI/TRIFrgCore.cpp:198-711 (region 0, pp ladder), :2389-2902 (1, ph ladder),
 * :1298-1811 (2, chalice), :1855-2368 (3, inverse chalice), :739-1252 (4, RPA; indices before the overlap's spin permutations).
 * Writes up to `capacity` rows {out channel, sign, first factor channel, second factor channel} and returns the number of
 * terms per buffer pair (256, or 64 for the RPA). Used by the parity tests to compare against the reference term by term. */
int pffrg_tri_terms(int region, int32_t *terms, int capacity);

/* Tables of the Gram form of the TRI RPA phase as the kernel walks them (rpaTriGram / triGramReduce, pffrg_kernels.cuh): the sum over
 * quadrature nodes and buffer pairs of R^{mu nu}[rid] = sum_i sum_k 2 eta(mu,k,nu) A^{p1 mu,p1 k}[rid1_i] B^{p2 k,p2 nu}[rid2_i]
 * (src/TRI/TRIFrgCore.cpp:733-1252) taken over the Gram blocks G^{(c1,c2)} = sum A^{c1} (x) B^{c2}. blocks[round * resident + slot] =
 * c1 | c2 << 4 (0xffff: none); per round one pass of words (slot * GBLK + rid1 * GS + rid2) | out << 13 | minus << 23 | multiplicity << 24
 * with out = (4 mu + nu) * n_sites + rid, GS = 8 ceil(n_sites / 8) + 1, GBLK = (GS - 1) GS; seg[2 * (round * warps + w)] = {begin, end}.
 * *rounds receives the number of rounds; returns the number of words. Host only; used by the CPU tests. */
int pffrg_trigram_tables(const pffrg_desc *desc, int resident, int warps, uint16_t *blocks, int block_capacity, uint32_t *terms, int capacity,
                         int32_t *seg, int seg_capacity, int32_t *rounds);

/* Device-internal order of the representative sites: order[k] = the reference's site index stored at position k of the device
 * layout. The two members of every pair {j, getInvertedSites()[j]} (src/Lattice.hpp:157-161) are neighbours, so that gathers with the
 * site-exchange flag (src/SU2/SU2VertexTwoParticle.hpp:369-387) touch the same cache lines as plain ones; site 0 stays first. Returns 1
 * if the order differs from the reference's, 0 if not. Host only; the order never shows at the boundary. */
int pffrg_site_order(const pffrg_desc *desc, int32_t *order /* [n_sites] */);

/* Term tables of the Gram form of the SU2 RPA lattice sum as the kernel walks them (rpaGram / gramReduce, pffrg_kernels.cuh): the sum
 * over the quadrature nodes of R[rid] = sum_i A[rid1_i] B[rid2_i] (Lattice::getOverlap(rid), src/Lattice.hpp:46-150, evaluated per node
 * at src/SU2/SU2FrgCore.cpp:250-266) is taken as sum_i G[rid1_i][rid2_i] over the Gram matrix G = sum_nodes A (x) B, rows worked off in
 * blocks of `rows_per_block`. Writes up to `capacity` 32-bit words ((rid1 - block * rows_per_block) * (Lp + 1) + rid2) | rid << 14 |
 * multiplicity << 22 | flush << 31, Lp = n_sites rounded up to 4. Whole rid lists are dealt to the warps; a warp's lists are padded to
 * groups of 4 words, concatenated, and lane l walks the groups [l T4, (l + 1) T4) serially (stored as [group][lane][4]); the last word
 * of a list's last group carries the flush flag. seg[2 * (block * warps + w)] = {first word, T4}; *conflict_degree (optional) = average
 * shared-memory bank-conflict degree of the walk (1 = conflict free). Returns the number of words. Host only; used by the CPU tests. */
int pffrg_gram_tables(const pffrg_desc *desc, int rows_per_block, int warps, uint32_t *terms, int capacity, int32_t *seg, double *conflict_degree);

/* measured FP64 multiply-add peak of a device in TFLOP/s (a 16-chain DFMA loop on every SM; ~10 ms): the denominator of the
 * FP64 roofline bench.py reports next to the HBM one. Returns a negative value on failure. */
double pffrg_fp64_peak(int device);
/* the same for the FP64 tensor-core path (mma.sync m8n8k4 f64), TFLOP/s; negative on failure */
double pffrg_dmma_peak(int device);

/* page-locked host memory for the arrays passed to set_state / get_state / get_flow (plain memory works too, but
 * transfers from pinned buffers run at full PCIe speed). Returns NULL on failure. */
void *pffrg_host_alloc(size_t bytes);
void pffrg_host_free(void *p);
/* page-lock memory the caller already owns (e.g. the `new float[]` vertex arrays of src/SU2/SU2VertexTwoParticle.hpp:76-77)
 * in place, so that set_state / get_state / get_flow on them run at full PCIe speed without a staging copy on the host. */
int pffrg_host_register(void *p, size_t bytes);
int pffrg_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif
