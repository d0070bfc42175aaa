/* pffrg_oracle.h -- CPU restatement ("port") of SpinParser's pf-FRG flow-equation hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under spinparser_b200/ may include, link or call this; it is the checker used
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs. The product path is the CUDA
 * library behind include/pffrg.h and fails loudly when that library is missing.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_port.py) against dumps of the UNMODIFIED reference
 * compiled in FP64 (oracle/_ref/oracle64, built by oracle/Makefile from /root/reference/src) that are committed
 * under tests/golden/, and through them against the reference's own golden files test/scripted/assets/test_reference{1,2,3}.ref.
 *
 * All arrays are in the reference's memory layout (file:line relative to /root/reference):
 *   v2   [Nw]                                    src/SU2/SU2VertexSingleParticle.hpp:100-102
 *   v4_c [su][t][rid], su = so(so+1)/2+uo        src/SU2/SU2VertexTwoParticle.hpp:595-605  (SU2: c = S,D; XYZ: c = X,Y,Z,D)
 *   v4   [su][t][mu][nu][rid]                    src/TRI/TRIVertexTwoParticle.hpp:65-71    (TRI: one array)
 */
#ifndef PFFRG_ORACLE_H
#define PFFRG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { PFO_SU2 = 0, PFO_XYZ = 1, PFO_TRI = 2 };

typedef struct pfo_problem
{
	int core;               /* PFO_SU2 | PFO_XYZ | PFO_TRI */
	int nw;                 /* number of positive mesh points */
	const double *mesh;     /* [nw] ascending positive frequencies (src/FrequencyDiscretization.hpp:174-192) */
	int L;                  /* number of representative sites (Lattice::size) */
	const int *sites_rid;   /* [L]   Lattice::getSites()          (src/Lattice.hpp:491) */
	const int *sites_perm;  /* [L*3] spin permutation of each */
	const int *inv_rid;     /* [L]   Lattice::getInvertedSites()  (src/Lattice.hpp:481) */
	const int *inv_perm;    /* [L*3] */
	const int *ov_off;      /* [L+1] CSR offsets of LatticeOverlap (src/Lattice.hpp:46-150) */
	const int *ov_rid1, *ov_rid2;   /* [ov_off[L]] */
	const int *ov_perm1, *ov_perm2; /* [ov_off[L]*3] transformed{X,Y,Z}{1,2} */
	int nrange;             /* sites in range of the reference site (Lattice::getRange(0)) */
	const int *rng_fwd_rid; /* [nrange] symmetryTransform(0, j)  */
	const int *rng_inv_rid; /* [nrange] symmetryTransform(j, 0)  */
	double spin_length;     /* SU2 only (src/SU2/SU2FrgCore.cpp:20-29) */
} pfo_problem;

/* number of v4 arrays in reference layout (2 / 4 / 1) and channels per site (2 / 4 / 16) */
int pfo_num_arrays(int core);
int pfo_num_channels(int core);

/* d/dLambda Sigma(w_i), i in [0,nw): {SU2,XYZ,TRI}FrgCore::_calculateVertexSingleParticle
 * (src/SU2/SU2FrgCore.cpp:139-169, src/XYZ/XYZFrgCore.cpp:164-193, src/TRI/TRIFrgCore.cpp:122-151). */
void pfo_v2_flow(const pfo_problem *p, double cutoff, const double *v2, const double *const *v4, double *v2flow);

/* d/dLambda Gamma for the work items n_items ids in items[] (NULL: 0..n_items-1), written to flow[c][item*L .. +L)
 * (TRI: flow[0][item*16L ..)): {SU2,XYZ,TRI}FrgCore::_calculateVertexTwoParticle
 * (src/SU2/SU2FrgCore.cpp:171-431, src/XYZ/XYZFrgCore.cpp:195-571, src/TRI/TRIFrgCore.cpp:153-3001).
 * OpenMP `parallel for schedule(guided)` over the items as in src/lib/LoadManager.hpp:551-557. */
void pfo_v4_flow(const pfo_problem *p, double cutoff, const double *v2, const double *v2flow, const double *const *v4,
                 const int *items, int n_items, double *const *flow);

/* Euler update x += (newCutoff - cutoff) * flow  (src/SU2/SU2FrgCore.cpp:111-134) */
void pfo_euler(double *x, const double *flow, long n, double cutoff, double new_cutoff);

/* number of kernel evaluations (quadrature nodes) of one channel whose transfer frequency is x */
int pfo_node_count(const pfo_problem *p, double cutoff, double x);

/* building blocks exported for the unit tests that mirror test/test_FrequencyDiscretization.cpp and test/test_Integrator.cpp */
int pfo_mesh_lesser(int nw, const double *mesh, double w);   /* index relative to the first positive point, negative side: -(i+1) */
int pfo_mesh_greater(int nw, const double *mesh, double w);
int pfo_mesh_offset(int nw, const double *mesh, double w);
void pfo_mesh_interpolate(int nw, const double *mesh, double w, int *lower, int *upper, double *bias);
double pfo_mesh_value(int nw, const double *mesh, int index);
typedef double (*pfo_scalar_fn)(double w, void *ctx);
double pfo_integrate_left(int nw, const double *mesh, double min, int max_index, pfo_scalar_fn f, void *ctx);
double pfo_integrate_right(int nw, const double *mesh, int min_index, double max, pfo_scalar_fn f, void *ctx);
double pfo_integrate_both(int nw, const double *mesh, double min, double max, pfo_scalar_fn f, void *ctx);

#ifdef __cplusplus
}
#endif
#endif
