// pffrg_jit.hpp -- interface of the lattice-specialised RPA code generator (see pffrg_jit.cpp)
#pragma once

#include <string>
#include <vector>

namespace pffrg
{
	// one multiply-add of the RPA sum: out[o] += mult * A[a] * B[b]; a, b index the staged operands [channel][rid]
	struct RpaTerm { int out, a, b, mult; };

	struct RpaProgram
	{
		std::vector<RpaTerm> terms; // for ONE variant
		int nOutputs = 0;           // outputs per variant
		int variants = 1;           // lanes are split into `variants` groups running the same code on shifted operands (SU2: the two channels)
		int lanesPerVariant = 32;   // nodes per warp
		int nb = 32;                // nodes per RPA phase (the kernel's NBT)
		int warps = 8;              // warps taking part in the RPA phase (the first `warps` of the CTA)
		long operandStride = 33;    // doubles between consecutive operand indices in the staging area (NB + 1)
		long operandBOffset = 0;    // doubles from operand A's staging buffer to operand B's
		long variantOperandStride = 0, variantOutputStride = 0;
		long outputCopyStride = 0;  // doubles between the two output copies of the kernel (C * L)
		int maxAccumulators = 8;    // outputs accumulated in registers at a time
		int chunk = 32;             // B operands cached in registers at a time
		int prefetch = 8;           // A operands in flight (software pipeline depth of the generated code)
		int cluster = 1;            // CTAs per thread-block cluster (template argument of clusterRendezvous in the generated code)
		bool resync = false;        // thread-block clusters: rendezvous again after every sub-tile of outputs (the CTAs of a cluster drift apart inside a long stream)
	};

	// CUDA source of `__device__ void pffrg::rpaSpecialised(int warp, int lane, int nb, const double *st, double *rpaOut)`
	std::string generateRpaSource(const RpaProgram &program);

	struct KernelSizes { int L, Lp, RL, nw; }; // baked into the run-time compiled kernel as constants

	// compile the vertex-flow kernel (embedded source + the generated RPA function) for sm_100a; returns an empty string on
	// success and the compiler log otherwise. The kernel is `pffrg_v4flow_jit` with v4FlowKernel's parameter list; `nbt` = nodes
	// staged per RPA phase by each of the `subs` sub-CTAs of a CTA, `threads` = threads of the whole CTA, `cluster` = CTAs per thread-block cluster (they rendezvous before every RPA phase).
	// `defines`: further preprocessor definitions (one per line), e.g. the PFFRG_GRAM_* set that selects the Gram form of the RPA phase
	// (then rpaSource is empty).
	std::string compileFlowKernel(int core, int nb, int nbt, int subs, int cluster, int threads, int minBlocks, const KernelSizes &sizes, const std::string &rpaSource, std::vector<char> &cubin, const std::string &defines = std::string(), std::string *cacheHit = nullptr);
}
