// pffrg.cu -- host side of libpffrg: the C ABI of include/pffrg.h on top of the kernels in pffrg_kernels.cuh.
//
// One handle = one GPU. The handle owns the device copies of the problem tables, the vertex state (v2, v4), the flow of
// the last step, and the per-step quadrature node table. Multi-GPU runs use one process per GPU: work items are split
// into contiguous cost-balanced ranges (replacing the dynamic master/worker chunking of src/lib/LoadManager.hpp:796-852),
// and after the Euler update every rank broadcasts its slice of the updated vertex (ncclBroadcast group on the compute
// stream; replaces the MPI_Bcast of src/lib/LoadManager.hpp:240 issued from src/SU2/SU2FrgCore.cpp:136).
#include "pffrg.h"
#include "pffrg_kernels.cuh"
#include "pffrg_measure.cuh"
#include "pffrg_jit.hpp"

#include <nccl.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <unistd.h>
#include <map>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

using namespace pffrg;

namespace
{
	thread_local std::string g_lastError = "no error";

	// NCCL is bound at run time, on the first multi-GPU call: a single-GPU core needs no NCCL at all, and a host process
	// that already carries an NCCL (e.g. the one bundled with PyTorch) shares it instead of loading a second copy.
	// PFFRG_NCCL_LIB overrides the library name.
	struct NcclApi
	{
		ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
		ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
		ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
		ncclResult_t (*GroupStart)() = nullptr;
		ncclResult_t (*GroupEnd)() = nullptr;
		ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
		ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
		ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
		const char *(*GetErrorString)(ncclResult_t) = nullptr;
		std::string error;
		bool ok = false;
	};
	NcclApi &nccl()
	{
		static NcclApi api = [] {
			NcclApi a;
			const char *name = getenv("PFFRG_NCCL_LIB");
			void *lib = dlopen(name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
			if (!lib) { a.error = std::string("cannot load NCCL: ") + dlerror(); return a; }
			auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p && a.error.empty()) a.error = std::string("NCCL symbol missing: ") + n; return p; };
			a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
			a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
			a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
			a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
			a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
			a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
			a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
			a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
			a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
			a.ok = a.error.empty();
			return a;
		}();
		return api;
	}

	int fail(int code, const char *fmt, ...)
	{
		char buf[1024];
		va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
		g_lastError = buf;
		return code;
	}

#define CUDA_TRY(expr)                                                                                                   \
	do {                                                                                                                 \
		cudaError_t e_ = (expr);                                                                                         \
		if (e_ != cudaSuccess) return fail(PFFRG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)
#define NCCL_TRY(expr)                                                                                                   \
	do {                                                                                                                 \
		ncclResult_t r_ = (expr);                                                                                        \
		if (r_ != ncclSuccess) return fail(PFFRG_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r_), __FILE__, __LINE__); \
	} while (0)

	template <typename T> struct DeviceArray
	{
		T *p = nullptr; size_t n = 0;
		cudaError_t alloc(size_t count) { release(); n = count; return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)); }
		cudaError_t upload(const std::vector<T> &h) { cudaError_t e = alloc(h.size()); if (e != cudaSuccess || h.empty()) return e; return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); }
		void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	};

	// per-core constants of the cost / byte / flop model
	struct CoreModel { int C, arrays, ladderTerms, localTerms, rpaTerms; };
	CoreModel modelOf(int core)
	{
		if (core == SU2) return { 2, 2, 10, 16, 2 };
		if (core == XYZ) return { 4, 4, 32, 64, 4 };
		return { 16, 1, 512, 1024, 128 };
	}
}

struct pffrg_context
{
	int core = 0, nw = 0, L = 0, Lp = 0, RL = 0, C = 0, nArrays = 0, device = 0;
	int64_t nf = 0;
	int64_t overlapTotal = 0, uniquePairs = 0;
	double spin = 0.5;
	std::vector<double> mesh;

	// device tables
	DeviceArray<double> dMesh;
	DeviceArray<int> dSitesRid, dInvRid, dSitesPerm, dInvPerm, dRngFwd, dRngInv, dSlotOff;
	DeviceArray<int4> dTasks;
	DeviceArray<unsigned> dWords;
	DeviceArray<unsigned> dGramTerms; DeviceArray<int> dGramSeg; // Gram form of the RPA phase (rpaGram), when selected
	int gramRows = 0;                                                 // rows per Gram block (0: not in use); TRI Gram form: resident channel-pair blocks
	int splitGather = 0, producerWarps = 0;                           // warp-specialised kernel: gather threads (0: not split); producer warps
	bool persistent = true;
	int smCount = 0;                                                  // SMs of the device (grid of the persistent warp-specialised kernel)
	DeviceArray<unsigned short> dTriBlocks; int triRounds = 0;         // TRI Gram form: channel pairs per (round, slot)
	int64_t triBlockCount = 0, gramWords = 0;                          // Gram forms: needed channel-pair blocks (TRI), words walked per RPA phase
	int itemOrder = 0;                                                // FlowConfig::order of the run-time compiled kernel (PFFRG_ORDER=t: t-major)
	std::vector<int> siteNewOf, siteOrder;                            // device-internal site order (RelabelledDesc); empty: the reference order
	DeviceArray<int> dSiteNewOf;
	DeviceArray<unsigned short> dMeshStart; int meshShift = 52, meshKeyBase = 0, meshKeys = 0;
	// lattice-specialised flow kernel (NVRTC), see pffrg_jit.cpp
	cudaLibrary_t jitLibrary = nullptr;
	cudaKernel_t jitKernel = nullptr;
	double jitCompileMs = 0.0;
	int autotuned = 0; // number of launch shapes that were timed at creation
	// state
	DeviceArray<double> dV4, dFlow4, dV2, dFlow2, dCutoff;
	DeviceArray<int> dCount; DeviceArray<double> dNodeW, dNodeWt;
	DeviceArray<int> dNan;
	DeviceArray<double> dStaging; // one reference-layout array, reused by set_state / get_state / get_flow
	DeviceArray<double> dChiPartial, dChi; DeviceArray<int> dChiCount; // correlation measurement (allocated on first use)
	int *hNan = nullptr;
	int nodeStride = 0;

	// launch configuration of the flow kernel
	int nb = 32, nbt = 32, rpaWarps = 8, minBlocks = 2, groups = 1, stride = 1, threads = 32, nslots = 1, subs = 1, cluster = 1; size_t smemBytes = 0;
	// (run-time compiled kernel: `threads` = all threads of a CTA = subs sub-CTAs of groups * stride threads (rounded to warps), nbt = nodes per RPA phase and sub-CTA)

	cudaStream_t stream = nullptr;
	cudaEvent_t ev[8] = {};
	bool haveState = false, haveFlow = false, flowGathered = true;
	double cutoff = 0.0;

	// partition
	int rank = 0, nRanks = 1;
	ncclComm_t comm = nullptr;
	std::vector<int64_t> bounds; // [nRanks + 1] item boundaries of the current step
	std::vector<double> rankTimes; // flow-kernel time of every rank in the last step (feedback of the partition), empty before the first
	DeviceArray<double> dTimes; double *hTimes = nullptr; bool balance = true;
	// exchange over peer memory (see eulerPushKernel): second state buffer, the mapped buffers of all ranks, arrival counters
	DeviceArray<double> dV4b; int cur = 0;           // cur: which of dV4 / dV4b holds the current state
	DeviceArray<SyncBlock> dSync; unsigned long long epoch = 0;
	bool p2p = false;
	double *peerV4[2][MAX_RANKS] = {}; SyncBlock *peerSync[MAX_RANKS] = {}; void *ipcOpened[3 * MAX_RANKS] = {}; int nIpcOpened = 0;
	int *hFlags = nullptr;                           // pinned: [0] a peer did not arrive in time
	bool timesPending = false;                       // hTimes is being filled by the last finalize_step
	bool finalizePending = false;                    // the last finalize_step has not been synchronised with the host yet
	DeviceArray<double> dVecStaging;                 // self-energy transfers in the caller's precision
	double *v4cur() const { return cur ? dV4b.p : dV4.p; }
	double *v4next() const { return cur ? dV4.p : dV4b.p; }
	int64_t userBegin = 0, userEnd = 0;
	int64_t curBegin = 0, curEnd = 0;

	pffrg_stats stats = {};

	Problem problem() const
	{
		Problem P;
		P.nw = nw; P.L = L; P.Lp = Lp; P.RL = RL; P.nf = (int)nf;
		P.mesh = dMesh.p; P.sites_rid = dSitesRid.p; P.inv_rid = dInvRid.p; P.sites_perm = dSitesPerm.p; P.inv_perm = dInvPerm.p;
		P.rpa_tasks = dTasks.p; P.rpa_slot_off = dSlotOff.p; P.rpa_words = dWords.p;
		P.gram_terms = dGramTerms.p; P.gram_seg = reinterpret_cast<const int2 *>(dGramSeg.p);
		P.trigram_blocks = dTriBlocks.p; P.trigram_rounds = triRounds;
		P.nrange = (int)dRngFwd.n; P.rng_fwd = dRngFwd.p; P.rng_inv = dRngInv.p; P.spin = spin;
		P.meshIndex.start = dMeshStart.p; P.meshIndex.shift = meshShift; P.meshIndex.keyBase = meshKeyBase; P.meshIndex.nKeys = meshKeys;
		return P;
	}
	NodeTable nodeTable() const { NodeTable N; N.count = dCount.p; N.wp = dNodeW.p; N.wt = dNodeWt.p; N.stride = nodeStride; return N; }
	size_t v4Elements() const { return (size_t)nf * RL; }
};

namespace
{
	int packPerm(const int32_t *p) { return (p[0] & 3) | ((p[1] & 3) << 2) | ((p[2] & 3) << 4); }

	// sites per channel run in the device layout: L rounded up so that every run starts on a 32-byte sector
	// (PFFRG_SITE_ALIGN overrides the granularity in doubles, for layout experiments)
	int paddedSites(int L)
	{
		int align = 4;
		if (const char *e = getenv("PFFRG_SITE_ALIGN")) align = std::max(1, atoi(e));
		return (L + align - 1) / align * align;
	}

	// Bucket index over the mesh (MeshIndex, pffrg_device.cuh): buckets = runs of equal leading bits of the IEEE representation,
	// as fine as possible with at most 2048 buckets between the first and the last mesh value.
	std::vector<unsigned short> buildMeshIndex(const std::vector<double> &mesh, int &shift, int &keyBase, int &nKeys)
	{
		auto bits = [](double x) { long long b; memcpy(&b, &x, 8); return b; };
		auto value = [](long long b) { double x; memcpy(&x, &b, 8); return x; };
		const int nw = (int)mesh.size();
		for (shift = 52 - 5; ; ++shift)
		{
			keyBase = (int)(bits(mesh.front()) >> shift);
			nKeys = (int)(bits(mesh.back()) >> shift) - keyBase + 1;
			if (nKeys <= 2048) break;
		}
		std::vector<unsigned short> start(nKeys + 1);
		for (int k = 0; k <= nKeys; ++k) start[k] = (unsigned short)firstGreater(mesh.data(), nw, value((long long)(keyBase + k) << shift));
		return start;
	}

	template <int CORE, int NB> size_t flowSmemBytes(int nw, int L, int groups, int nbt, int subs) { return FlowSmem<CORE, NB>(nw, L, groups, nbt, subs).total; }
	// Operand geometry of the Gram form: SU2 -- the lattice's sites; XYZ -- the three spin channels of a site as virtual sites c L + j (the
	// density channel rides in the second component of the c = 0 entries), outputs o = c L + rid and 3 L + rid (see gramcfg, pffrg_kernels.cuh)
	struct GramGeometry { int lv, lp, lout; };
	GramGeometry gramGeometry(int core, int L, int Lp)
	{
		if (core == XYZ) return { 3 * L, (3 * L + 3) / 4 * 4, 4 * L };
		return { L, Lp, L };
	}
	// shared memory of the kernel with the Gram form of the RPA phase: gramRows rows of the Gram matrix next to nbt staged nodes (Lp: padded operand sites)
	size_t gramSmemBytes(int nb, int nw, int L, int Lp, int groups, int nbt, int gramRows, int tableCopies = 1, int core = SU2)
	{
		if (core == XYZ) return nb == 32 ? FlowSmem<XYZ, 32>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total : nb == 16 ? FlowSmem<XYZ, 16>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total
		                                : FlowSmem<XYZ, 8>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total;
		return nb == 32 ? FlowSmem<SU2, 32>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total : nb == 16 ? FlowSmem<SU2, 16>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total
		                : FlowSmem<SU2, 8>(nw, L, groups, nbt, 1, gramRows, Lp, tableCopies).total;
	}
	// nbt = nodes staged per RPA phase (0: same as the gather batch nb)
	size_t flowSmemBytes(int core, int nb, int nw, int L, int groups, int nbt = 0, int subs = 1)
	{
		if (nbt <= 0) nbt = nb;
		if (core == SU2) return nb == 32 ? flowSmemBytes<SU2, 32>(nw, L, groups, nbt, subs) : nb == 16 ? flowSmemBytes<SU2, 16>(nw, L, groups, nbt, subs) : flowSmemBytes<SU2, 8>(nw, L, groups, nbt, subs);
		if (core == XYZ) return nb == 32 ? flowSmemBytes<XYZ, 32>(nw, L, groups, nbt, subs) : nb == 16 ? flowSmemBytes<XYZ, 16>(nw, L, groups, nbt, subs) : flowSmemBytes<XYZ, 8>(nw, L, groups, nbt, subs);
		return nb == 32 ? flowSmemBytes<TRI, 32>(nw, L, groups, nbt, subs) : nb == 16 ? flowSmemBytes<TRI, 16>(nw, L, groups, nbt, subs) : nb == 8 ? flowSmemBytes<TRI, 8>(nw, L, groups, nbt, subs) : flowSmemBytes<TRI, 4>(nw, L, groups, nbt, subs);
	}

	// Launch shape of the lattice-specialised kernel. The RPA phase runs as ONE instruction stream shared by `nodeGroups`
	// warps (one per SM sub-partition when there are four), each on its own group of 16 (SU2) or 32 staged nodes, so the
	// number of staged nodes nbt = nodeGroups * lanes decides the shared-memory footprint. Environment overrides for tuning runs:
	// PFFRG_JIT_NB, PFFRG_JIT_NBT, PFFRG_JIT_TILES, PFFRG_JIT_MINBLOCKS.
	struct JitShape { int nb, nbt, rpaWarps, minBlocks; size_t smem; int subs = 1; int cluster = 1; int gramRows = 0, gramThreads = 0; int producer = 0; /* producer warps */ int splitGather = 0; /* warp-specialised kernel: gather threads (0: not split) */ int regsGather = 0, regsRpa = 0, regsProducer = 0, regsLaunch = 0; int tableCopies = 1; /* access-buffer table blocks (producer kernels: 2..4) */ };

	// tiles of the busiest warp when nw warps share the rt x ct 8x8 tiles of a Gram block (gramcfg::bestRowWarps, pffrg_kernels.cuh)
	int gramBusiestTiles(int rt, int ct, int nw)
	{
		int best = 1 << 30;
		for (int wp = 1; wp <= nw; ++wp) if (nw % wp == 0) best = std::min(best, ((rt + wp - 1) / wp) * ((ct + nw / wp - 1) / (nw / wp)));
		return best;
	}

	// Launch shape of the kernel with the Gram form of the RPA phase (rpaGram, pffrg_kernels.cuh): the staged nodes (nbt, as many as fit:
	// the overlap list is walked once per RPA phase) and one block of PB rows of the Gram matrix (a multiple of 8, at most 64, at most 16
	// accumulator tiles per warp) share the shared memory. Two CTAs per SM where 32 staged nodes still fit. Environment overrides:
	// PFFRG_JIT_NB, PFFRG_JIT_NBT, PFFRG_JIT_MINBLOCKS, PFFRG_GRAM_PB.
	// `threads` = worker threads; producer: one more warp builds the access buffers a batch ahead (two table blocks), one CTA per SM
	// tableCopies: access-buffer table blocks (0: one, or two with producer warps; -1: the warp-specialised kernel -- two or three blocks and
	// gather batches of up to 32 nodes, rated below)
	JitShape chooseGramShape(int nw, int L, int Lp, int groups, int threads, size_t smemMax, int64_t uniquePairs, int producer = 0, int maxCtas = 2, int maxTiles = 16, int tableCopies = 0, int core = SU2)
	{
		JitShape best = { 0, 0, 0, 0, 0 };
		const bool split = tableCopies < 0;
		if (tableCopies == 0) tableCopies = producer ? 2 : 1;
		const int gemmThreads = threads / 32 * 32, warps = gemmThreads / 32, ct = (Lp + 7) / 8;
		if (warps < 1) return best;
		int forcedNb = 0, forcedNbt = 0, forcedCtas = 0, forcedPb = 0, forcedTables = 0;
		if (const char *e = getenv("PFFRG_JIT_NB")) forcedNb = atoi(e);
		if (const char *e = getenv("PFFRG_JIT_NBT")) forcedNbt = atoi(e);
		if (const char *e = getenv("PFFRG_JIT_MINBLOCKS")) forcedCtas = std::max(1, atoi(e));
		if (const char *e = getenv("PFFRG_GRAM_PB")) forcedPb = std::max(8, atoi(e) / 8 * 8);
		if (const char *e = getenv("PFFRG_SPLIT_TABLES")) forcedTables = std::min(4, std::max(2, atoi(e)));
		const int pbMax = std::min(64, (Lp + 7) / 8 * 8);
		const int nbts[] = { 64, 48, 32, 24, 16, 8 };
		// Every shape that fits is rated with a coarse model of what the choice costs per work item (clocks; ~64 t-channel nodes per item):
		// every RPA phase walks the term array once (~60 clocks per group of 128 words and warp) and pays two barriers and a store of the
		// block per row block (~1500 clocks); small gather batches cost gather throughput (measured: pyrochlore-r8 +15 % with batches of 8).
		// Warp-specialised kernel (measured on B200): batches of 32 nodes instead of 16 -- pyrochlore-r8 70.3 -> 68.4 ms, cubic-r7 16.3 -> 15.6 ms;
		// a third table block -- cubic-r7 15.6 -> 15.0 ms, nothing on pyrochlore-r8 (where it would cost rows of the Gram block).
		double bestCost = 0.0;
		for (int ctas = 2; ctas >= 1; --ctas)
		{
			if ((forcedCtas && ctas != forcedCtas) || ctas > maxCtas) continue;
			if (!forcedCtas && ctas == 2 && threads > 256) continue; // the block update needs more than 64 registers per thread

			const size_t budget = ctas >= 2 ? (smemMax + 1024) / ctas - 1024 : smemMax;
			for (int nbt : nbts)
			{
				if (forcedNbt ? nbt != forcedNbt : (ctas == 2 && nbt < 32)) continue;
				for (int nb : { 32, 16, 8 })
				{
					if (nbt % nb || (forcedNb ? nb != forcedNb : (nb == 32 && !split))) continue; // (batches of 32 nodes: warp-specialised kernel, or PFFRG_JIT_NB=32)
					// table blocks: the warp-specialised kernel tries three, then two (PFFRG_SPLIT_TABLES = 2..4 forces one value); the others have a fixed number
					const int firstCopies = !split ? tableCopies : forcedTables ? forcedTables : 3, lastCopies = !split ? tableCopies : forcedTables ? forcedTables : 2;
					for (int copies = firstCopies; copies >= lastCopies; --copies)
					{
						for (int pb = pbMax; pb >= 8; pb -= 8)
						{
							if (forcedPb && pb != std::min(forcedPb, pbMax)) continue;
							if ((long)pb * (Lp + 1) > (1l << 14)) continue;           // a term word addresses the Gram block with 14 bits
							const int blocks = (Lp + pb - 1) / pb, lastRt = (Lp - (blocks - 1) * pb + 7) / 8;
							if (gramBusiestTiles(pb / 8, ct, warps) > maxTiles || gramBusiestTiles(lastRt, ct, warps) > maxTiles) continue; // (8 accumulator registers per tile)
							const size_t smem = gramSmemBytes(nb, nw, L, Lp, groups, nbt, pb, copies, core);
							if (smem > budget) continue;
							const double phases = (64 + nbt - 1) / nbt;
							double cost = phases * (1.03 * (double)uniquePairs / 128.0 / warps * 60.0 + blocks * 1500.0) + (nb == 8 ? (producer ? 6000.0 : 15000.0) : 0.0); // (with producer warps the access-buffer phases are off the critical path; measured +4 % with batches of 8)
							if (split) cost += (nb == 16 ? 3000.0 : 0.0) + (copies == 2 ? 1000.0 : 0.0);
							if (ctas == 2) cost *= 0.8; // two resident CTAs overlap their phases
							if (!best.nb || cost < bestCost) { best = { nb, nbt, threads / 32, ctas, smem }; best.gramRows = pb; best.gramThreads = gemmThreads; best.producer = producer; best.tableCopies = copies; bestCost = cost; }
						}
					}
				}
			}
		}
		return best;
	}

	void buildTriGramTables(const pffrg_desc *d, int L, int resident, int warps, std::vector<unsigned short> &blocks, std::vector<unsigned> &terms, std::vector<int> &seg, double *conflictDegree = nullptr);
	// Term tables of rpaGram / gramReduce (pffrg_kernels.cuh), see the definition below
	void buildGramTables(const pffrg_desc *d, int L, int Lp, int PB, int warps, std::vector<unsigned> &terms, std::vector<int> &seg, double *conflictDegree = nullptr);
	JitShape chooseJitShape(int core, int nw, int L, int groups, int warps, size_t smemMax)
	{
		const int lanes = core == SU2 ? 16 : 32;
		JitShape best = { 0, 0, 0, 0, 0 };
		const size_t half = (smemMax + 1024) / 2 - 1024; // two CTAs per SM (1 KB per CTA is reserved by the driver)
		// (node groups, CTAs per SM) in order of preference (measured on B200: two resident CTAs beat a larger RPA batch)
		const int order[6][2] = { { 4, 2 }, { 2, 2 }, { 4, 1 }, { 2, 1 }, { 1, 2 }, { 1, 1 } };
		for (int pass = 0; pass < 6 && !best.nb; ++pass)
		{
			const int nodeGroups = order[pass][0], ctas = order[pass][1];
			if (warps < nodeGroups) continue; // every node group needs a warp of its own (rpaSpecialised: grp = warp % nodeGroups)
			const int nbt = nodeGroups * lanes;
			// two output tiles (instruction streams) per node group, at least four RPA warps. Measured: pyrochlore-r8 (4 groups)
			// 316 ms with one tile, 272 ms with two; honeycomb-r7 XYZ (2 groups) 32.7 ms with two tiles, 43.1 ms with four
			const int rpaWarps = std::min(warps, std::max(4, 2 * nodeGroups)) / nodeGroups * nodeGroups;
			// a smaller gather batch shrinks the access-buffer tables when the staged operands leave little room
			for (int nb = std::min(32, nbt); nb >= 8 && !best.nb; nb >>= 1)
			{
				const size_t smem = flowSmemBytes(core, nb, nw, L, groups, nbt);
				if (smem <= (ctas == 2 ? half : smemMax)) best = { nb, nbt, rpaWarps, ctas, smem };
			}
		}
		if (const char *e = getenv("PFFRG_JIT_NBT"))
		{
			const int nbt = std::max(lanes, atoi(e) / lanes * lanes);
			int nb = std::min(32, nbt);
			if (const char *f = getenv("PFFRG_JIT_NB")) nb = atoi(f);
			if ((nb == 8 || nb == 16 || nb == 32) && nbt % nb == 0 && warps >= nbt / lanes)
			{
				const size_t smem = flowSmemBytes(core, nb, nw, L, groups, nbt);
				if (smem <= smemMax) best = { nb, nbt, std::min(warps, std::max(4, 2 * (nbt / lanes))) / (nbt / lanes) * (nbt / lanes), smem <= half ? 2 : 1, smem };
			}
		}
		if (best.nb)
		{
			if (const char *e = getenv("PFFRG_JIT_TILES")) best.rpaWarps = std::min(warps, std::max(1, atoi(e)) * (best.nbt / lanes));
			if (const char *e = getenv("PFFRG_JIT_MINBLOCKS")) best.minBlocks = std::max(1, atoi(e));
		}
		return best;
	}

	template <int CORE, int NB>
	cudaError_t launchFlow(pffrg_context *h, int64_t begin, int64_t count)
	{
		auto kernel = v4FlowKernel<CORE, NB>;
		if (h->threads > 256) return cudaErrorInvalidConfiguration; // __launch_bounds__(256) of the precompiled kernels
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes);
		if (e != cudaSuccess) return e;
		FlowConfig cfg; cfg.groups = h->groups; cfg.stride = h->stride; cfg.nslots = h->nslots; cfg.smemBytes = (int)h->smemBytes; cfg.items = (int)count; cfg.order = 0;
		kernel<<<(unsigned)count, h->threads, h->smemBytes, h->stream>>>(h->problem(), h->nodeTable(), cfg, h->v4cur(), h->dFlow4.p, (int)begin, h->dNan.p);
		return cudaGetLastError();
	}

	RpaProgram buildRpaProgram(const pffrg_desc *d, int core, int nbt, int rpaWarps);

	// threads of a CTA: `groups` groups of `stride` threads, one group per quadrature node at a time (used by pffrg_create and
	// by the device-less pffrg_jit_compile_check, which must arrive at the same kernel)
	struct LaunchGeometry { int stride, groups, threads; };
	LaunchGeometry chooseGeometry(int L, bool runtimeCompiled)
	{
		LaunchGeometry g;
		const int padded = (L + 31) / 32 * 32;
		g.stride = (padded - L) * 4 <= padded ? padded : L;
		if (const char *e = getenv("PFFRG_PAD_GROUPS")) g.stride = atoi(e) ? padded : L; // tuning override
		int threadTarget = (runtimeCompiled && g.stride > 128) ? 384 : 256; // two groups (nodes in flight) also for L > 128; PFFRG_THREADS: tuning override (values above 256 only work with the run-time compiled kernel)
		if (const char *e = getenv("PFFRG_THREADS")) threadTarget = std::min(1024, std::max(64, atoi(e)));
		g.groups = std::max(1, threadTarget / g.stride);
		g.threads = std::max(64, (g.groups * g.stride + 31) / 32 * 32);
		return g;
	}

	// outputs accumulated in registers at a time: 16 when the kernel has the whole register file of an SM for 256 threads (255
	// registers per thread; pyrochlore-r8: 273 -> 241 ms, fewer operand loads and a shorter instruction stream), else 8
	int defaultAccumulators(int threads, int minBlocks) { return threads * minBlocks <= 256 ? 16 : 8; }

	// tuning knobs of the generated RPA code (defaults chosen on B200)
	void applyJitKnobs(RpaProgram &prog)
	{
		if (const char *e = getenv("PFFRG_JIT_CHUNK")) prog.chunk = std::max(4, atoi(e));
		if (const char *e = getenv("PFFRG_JIT_ACC")) prog.maxAccumulators = std::max(1, atoi(e));
		if (const char *e = getenv("PFFRG_JIT_PREFETCH")) prog.prefetch = std::max(1, atoi(e));
		if (const char *e = getenv("PFFRG_JIT_RESYNC")) prog.resync = atoi(e) != 0;
	}

	cudaError_t launchFlowDispatch(pffrg_context *h, int64_t begin, int64_t count)
	{
		if (count <= 0) return cudaSuccess;
		if (h->jitKernel)
		{
			Problem P = h->problem(); NodeTable N = h->nodeTable();
			FlowConfig cfg; cfg.groups = h->groups; cfg.stride = h->stride; cfg.nslots = h->nslots; cfg.smemBytes = (int)h->smemBytes; cfg.items = (int)count;
			cfg.order = (h->subs == 1 && h->cluster == 1) ? h->itemOrder : 0;
			const double *v4 = h->v4cur(); double *flow = h->dFlow4.p; int itemBegin = (int)begin; int *nan = h->dNan.p;
			void *args[] = { &P, &N, &cfg, &v4, &flow, &itemBegin, &nan };
			int64_t ctas = (count + h->subs - 1) / h->subs; // padded to whole clusters (the kernel carries __cluster_dims__)
			// warp-specialised kernel: persistent CTAs, one per SM, each working on every ctas-th item (v4FlowBodySplit); PFFRG_PERSISTENT=0: one CTA per item
			if (h->splitGather > 0 && h->persistent && h->smCount > 0) ctas = std::min<int64_t>(ctas, h->smCount);
			if (cfg.order == 1) ctas = ((begin + count - 1) / h->nw - begin / h->nw + 1) * h->nw; // t-major: whole (s,u) blocks
			return cudaLaunchKernel((const void *)h->jitKernel, dim3((unsigned)((ctas + h->cluster - 1) / h->cluster * h->cluster)), dim3(h->threads), args, h->smemBytes, h->stream);
		}
		if (h->core == SU2) return h->nb == 32 ? launchFlow<SU2, 32>(h, begin, count) : h->nb == 16 ? launchFlow<SU2, 16>(h, begin, count) : launchFlow<SU2, 8>(h, begin, count);
		if (h->core == XYZ) return h->nb == 32 ? launchFlow<XYZ, 32>(h, begin, count) : h->nb == 16 ? launchFlow<XYZ, 16>(h, begin, count) : launchFlow<XYZ, 8>(h, begin, count);
		return h->nb == 32 ? launchFlow<TRI, 32>(h, begin, count) : h->nb == 16 ? launchFlow<TRI, 16>(h, begin, count) : h->nb == 8 ? launchFlow<TRI, 8>(h, begin, count) : launchFlow<TRI, 4>(h, begin, count);
	}

	// Compile and load the lattice-specialised kernel. Controlled by the environment: PFFRG_JIT=0 disables it,
	// PFFRG_JIT_MAX_TERMS (default 60000) bounds the straight-line code size (compile time grows with it).
	float elapsed(cudaEvent_t a, cudaEvent_t b);

	// one launch shape of the run-time compiled kernel
	struct JitCandidate { int threads, groups; JitShape shape; cudaLibrary_t library; cudaKernel_t kernel; float ms; };

	// four node groups are out of reach for small CTAs; this shape runs two node groups in up to four CTAs of `warps` warps per SM
	JitShape smallCtaShape(int core, int nw, int L, int groups, int warps, size_t smemMax)
	{
		const int lanes = core == SU2 ? 16 : 32, nbt = 2 * lanes, nb = 16;
		const size_t smem = flowSmemBytes(core, nb, nw, L, groups, nbt);
		const size_t quarter = (smemMax + 1024) / 4 - 1024;
		if (smem > quarter) return { 0, 0, 0, 0, 0 };
		return { nb, nbt, std::min(warps, 4), 4, smem };
	}

	// A CTA of `subs` sub-CTAs (one work item each, `warps` warps each) that share ONE RPA phase over subs * nbt staged nodes:
	// the straight-line RPA code is streamed through the instruction caches once per `subs` items. nbt / nb are per sub-CTA.
	JitShape subCtaShape(int core, int nw, int L, int groups, int warps, int subs, int nbt, int nb, int ctas, size_t smemMax)
	{
		const int lanes = core == SU2 ? 16 : 32;
		if (subs < 2 || subs > 4 || nbt % lanes || nbt % nb || warps * subs > 32) return { 0, 0, 0, 0, 0 };
		const int nodeGroups = subs * nbt / lanes, total = warps * subs;
		if (total < nodeGroups) return { 0, 0, 0, 0, 0 }; // every node group needs a warp
		const size_t smem = flowSmemBytes(core, nb, nw, L, groups, nbt, subs);
		if (smem > (smemMax + 1024) / ctas - 1024) return { 0, 0, 0, 0, 0 };
		JitShape s = { nb, nbt, total / nodeGroups * nodeGroups, ctas, smem };
		s.subs = subs;
		return s;
	}

	bool wantGram(int core, int64_t uniquePairs)
	{
		// XYZ: warp-specialised kernel with the spin channels as virtual sites (at most 63 representative sites; a lattice that does not fit
		// stays on the straight-line code). Same threshold: honeycomb-r10 (3 323 merged terms) 79.2 -> 54.2 ms, honeycomb-r7 (944) 26.2 -> 26.1 ms
		if (core != SU2 && core != XYZ) return false;
		long minTerms = 2000; // (cubic-r7, 3453 terms: straight-line code 17.9 ms, warp-specialised Gram kernel 16.3 ms; square-r4, 136 terms: 0.69 / 1.01 ms)
		if (const char *e = getenv("PFFRG_GRAM_MIN_TERMS")) minTerms = atol(e);
		const char *form = getenv("PFFRG_RPA");
		return form ? std::string(form) == "gram" : uniquePairs > minTerms;
	}

	std::string gramDefines(const JitShape &s, int core = SU2, int L = 0, int Lp = 0)
	{
		const GramGeometry geo = gramGeometry(core, L, Lp);
		return (core == XYZ ? "#define PFFRG_GRAM_LP " + std::to_string(geo.lp) + "\n#define PFFRG_GRAM_LV " + std::to_string(geo.lv) + "\n#define PFFRG_GRAM_LOUT " + std::to_string(geo.lout) + "\n" : std::string()) +
		       "#define PFFRG_GRAM 1\n#define PFFRG_GRAM_THREADS " + std::to_string(s.gramThreads) + "\n#define PFFRG_GRAM_PB " + std::to_string(s.gramRows) +
		       "\n" + (s.producer ? "#define PFFRG_PRODUCER " + std::to_string(s.producer) + "\n" : std::string()) +
		       (s.splitGather ? "#define PFFRG_SPLIT 1\n#define PFFRG_SPLIT_GATHER_THREADS " + std::to_string(s.splitGather) + "\n#define PFFRG_SPLIT_REGS_GATHER " + std::to_string(s.regsGather) +
		                        "\n#define PFFRG_SPLIT_TABLES " + std::to_string(s.tableCopies) + "\n#define PFFRG_SPLIT_REGS_LAUNCH " + std::to_string(s.regsLaunch) + "\n#define PFFRG_SPLIT_REGS_RPA " + std::to_string(s.regsRpa) + "\n#define PFFRG_SPLIT_REGS_PRODUCER " + std::to_string(s.regsProducer) + "\n" : std::string());
	}
	// producer warp for the Gram kernel (v4FlowBodyProducer): opt-in with PFFRG_PRODUCER=1 while it is being measured
	int wantProducer() // number of producer warps (0: none)
	{
		const char *e = getenv("PFFRG_PRODUCER");
		return e ? std::min(4, std::max(0, atoi(e))) : 0;
	}

	// Warp-specialised Gram kernel (v4FlowBodySplit, pffrg_kernels.cuh): gather warps, RPA warps and producer warps as separate warp groups
	// of one CTA. Default for the lattices that get the Gram form by default (`large`: more than PFFRG_GRAM_MIN_TERMS merged overlap terms;
	// measured on B200: pyrochlore-r8 89.4 -> 72.2 ms, pyrochlore-r10 213 -> 194 ms; cubic-r7 with PFFRG_RPA=gram 19.2 -> 20.0 ms);
	// PFFRG_SPLIT=0 / 1 overrides. PFFRG_SPLIT_RPA_THREADS, PFFRG_SPLIT_REGS_RPA / _PRODUCER, PFFRG_PRODUCER (producer warps, default 4)
	// override the shape.
	bool wantSplit(bool large)
	{
		const char *e = getenv("PFFRG_SPLIT");
		return e ? atoi(e) != 0 : large;
	}

	// Launch of the SU2 kernel with the Gram form of the RPA phase: shape, threads per CTA, gather groups and the number of warps that walk the term array
	struct GramLaunch { JitShape shape; int threads; int reduceWarps; int groups; };
	GramLaunch chooseGramLaunch(int nw, int L, int Lp, int groups, int stride, int threads, size_t smemMax, int64_t uniquePairs, bool large, int core = SU2)
	{
		GramLaunch g = { { 0, 0, 0, 0, 0 }, threads, threads / 32, groups };
		const int workers = threads / 32 * 32;
		const GramGeometry geo = gramGeometry(core, L, Lp);
		if (core == XYZ && geo.lout > 255) return g; // an output index has 8 bits in a term word
		if (wantSplit(large) || core == XYZ)
		{
			// gather warps: at most 256 threads -- two nodes in flight up to 128 sites per group, one above (the register file is what bounds
			// the loads in flight); RPA warps: one warp group, two where the Gram matrix is large (L > 128)
			const int splitGroups = std::max(1, std::min(groups, 256 / stride));
			const int gather = (splitGroups * stride + 127) / 128 * 128;
			int rpa = (stride > 128 || geo.lp > 128) ? 256 : 128, producer = wantProducer() ? wantProducer() : 4;
			if (const char *e = getenv("PFFRG_SPLIT_RPA_THREADS")) rpa = std::max(128, atoi(e) / 128 * 128);
			const int total = gather + rpa + 128;
			if (total <= 1024)
			{
				// register file: what the launch allocates (registers per thread of the whole CTA, a multiple of 8) is re-partitioned
				const int pool = 65536 / total / 8 * 8 * total;
				int regsProducer = 40, regsRpa = rpa > 128 ? 120 : 200;
				if (const char *e = getenv("PFFRG_SPLIT_REGS_PRODUCER")) regsProducer = std::max(24, atoi(e) / 8 * 8);
				if (const char *e = getenv("PFFRG_SPLIT_REGS_RPA")) regsRpa = std::max(24, atoi(e) / 8 * 8);
				const int regsGather = std::min(256, (pool - 128 * regsProducer - rpa * regsRpa) / gather / 8 * 8);
				// one CTA per SM (the warp groups re-partition its whole register file); accumulator tiles of the block update as the RPA warps' registers allow
				JitShape s = chooseGramShape(nw, L, geo.lp, splitGroups, rpa, smemMax, core == XYZ ? 4 * uniquePairs : uniquePairs, producer, 1, std::min(16, (regsRpa - 56) / 8), -1, core);
				if (s.nb && regsGather >= 96)
				{
					s.splitGather = gather; s.regsLaunch = pool / total; s.regsGather = regsGather; s.regsRpa = regsRpa; s.regsProducer = regsProducer;
					g.shape = s; g.threads = total; g.reduceWarps = rpa / 32; g.groups = splitGroups;
					return g;
				}
			}
		}
		if (core == XYZ) return g; // the XYZ Gram form exists in the warp-specialised kernel only
		if (wantProducer() && workers + 32 * wantProducer() <= 1024)
		{
			g.shape = chooseGramShape(nw, L, Lp, groups, workers, smemMax, uniquePairs, wantProducer());
			if (g.shape.nb) { g.threads = workers + 32 * g.shape.producer; g.reduceWarps = workers / 32; return g; }
		}
		g.shape = chooseGramShape(nw, L, Lp, groups, threads, smemMax, uniquePairs);
		return g;
	}

	// TRI Gram form (rpaTriGram): gather batch = staged nodes = 8; as many resident channel-pair blocks as the shared memory holds
	// (offsets of a term word: 13 bits; at most 16 tiles per warp). PFFRG_TRIGRAM_RESIDENT overrides the number of blocks.
	JitShape chooseTriGramShape(int nw, int L, int groups, int threads, size_t smemMax)
	{
		JitShape best = { 0, 0, 0, 0, 0 };
		const int LT = (L + 7) / 8, LpT = 8 * LT, GBLK = LpT * (LpT + 1), warps = threads / 32;
		if (16 * L > 1024 || warps < 1) return best;
		int forced = 0;
		if (const char *e = getenv("PFFRG_TRIGRAM_RESIDENT")) forced = std::max(1, atoi(e));
		for (int resident = 16; resident >= 1 && !best.nb; --resident)
		{
			if (forced && resident != forced) continue;
			if (resident * GBLK > (1 << 13) || (resident * LT * LT + warps - 1) / warps > 16) continue;
			const size_t smem = FlowSmem<TRI, 8>(nw, L, groups, 8, 1, resident, LpT).total;
			if (smem <= smemMax) { best = { 8, 8, warps, 1, smem }; best.gramRows = resident; best.gramThreads = warps * 32; }
		}
		return best;
	}
	std::string triGramDefines(const JitShape &s)
	{
		return "#define PFFRG_TRIGRAM 1\n#define PFFRG_TRIGRAM_RESIDENT " + std::to_string(s.gramRows) + "\n#define PFFRG_GRAM_THREADS " + std::to_string(s.gramThreads) + "\n";
	}
	// Off by default: measured on B200 (kagome-DM r7, Nw 64) the Gram form takes 1249 ms per step against 993 ms of the table-driven phase.
	// A TRI node stages 20 KB of operands, so only 8 nodes (K = 16 with the two buffer pairs) fit per RPA phase next to 3 of the 96 needed
	// channel-pair blocks: 32 rounds of update -> store -> reduce per phase with 4 tensor-core steps each, and the per-round instruction
	// overhead (ncu: 790 warp instructions per 40 DMMA, 625 in the reduction) outweighs the saved multiply-adds. PFFRG_RPA=gram selects it.
	bool wantTriGram(int core)
	{
		if (core != TRI) return false;
		const char *form = getenv("PFFRG_RPA");
		return form && std::string(form) == "gram";
	}

	// compile (or take from the on-disk cache) and load one kernel. A cache entry the driver rejects -- truncated by a crash, written by another
	// toolkit -- is deleted and the kernel compiled afresh once.
	int compileAndLoad(pffrg_context *h, JitCandidate &c, int subs, int cluster, const std::string &rpaSource, const std::string &defines, const char *what)
	{
		for (int attempt = 0; attempt < 2; ++attempt)
		{
			std::vector<char> cubin; std::string cacheHit;
			const std::string err = compileFlowKernel(h->core, c.shape.nb, c.shape.nbt, subs, cluster, c.threads, c.shape.minBlocks, KernelSizes{ h->L, h->Lp, h->RL, h->nw }, rpaSource, cubin, defines, &cacheHit);
			if (!err.empty()) return fail(PFFRG_ERR_CUDA, "run-time compilation of %s failed: %s", what, err.c_str());
			c.library = nullptr;
			cudaError_t e = cudaLibraryLoadData(&c.library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
			if (e != cudaSuccess) c.library = nullptr;
			if (e == cudaSuccess) e = cudaLibraryGetKernel(&c.kernel, c.library, "pffrg_v4flow_jit");
			if (e == cudaSuccess) e = cudaFuncSetAttribute((const void *)c.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.shape.smem);
			if (e == cudaSuccess) return PFFRG_OK;
			cudaGetLastError();
			if (c.library) cudaLibraryUnload(c.library);
			c.library = nullptr; c.kernel = nullptr;
			if (attempt == 0 && !cacheHit.empty())
			{
				fprintf(stderr, "[pffrg] cached kernel %s rejected (%s): deleted, compiling afresh\n", cacheHit.c_str(), cudaGetErrorString(e));
				std::remove(cacheHit.c_str());
				continue;
			}
			return fail(PFFRG_ERR_CUDA, "loading %s failed: %s", what, cudaGetErrorString(e));
		}
		return fail(PFFRG_ERR_CUDA, "loading %s failed", what);
	}

	int compileCandidate(pffrg_context *h, const pffrg_desc *d, JitCandidate &c)
	{
		if (c.shape.gramRows > 0 && h->core == TRI) return compileAndLoad(h, c, 1, 1, std::string(), triGramDefines(c.shape), "the TRI flow kernel (Gram form)");
		if (c.shape.gramRows > 0) return compileAndLoad(h, c, 1, 1, std::string(), gramDefines(c.shape, h->core, h->L, h->Lp), "the flow kernel (Gram form)");
		RpaProgram prog = buildRpaProgram(d, h->core, c.shape.nbt * c.shape.subs, c.shape.rpaWarps);
		if (prog.warps < std::max(1, prog.nb / prog.lanesPerVariant) || prog.warps > c.threads / 32)
			return fail(PFFRG_ERR_STATE, "launch shape with %d RPA warps for %d node groups (%d threads): staged nodes would be dropped", prog.warps, prog.nb / prog.lanesPerVariant, c.threads);
		prog.maxAccumulators = defaultAccumulators(c.threads, c.shape.minBlocks);
		prog.cluster = c.shape.cluster;
		applyJitKnobs(prog);
		return compileAndLoad(h, c, c.shape.subs, c.shape.cluster, generateRpaSource(prog), std::string(), "the specialised flow kernel");
	}

	void adoptCandidate(pffrg_context *h, const JitCandidate &c)
	{
		h->jitLibrary = c.library; h->jitKernel = c.kernel;
		h->threads = c.threads; h->groups = c.groups;
		h->nb = c.shape.nb; h->nbt = c.shape.nbt; h->rpaWarps = c.shape.rpaWarps; h->minBlocks = c.shape.minBlocks; h->smemBytes = c.shape.smem; h->subs = c.shape.subs; h->cluster = c.shape.cluster;
		h->gramRows = c.shape.gramRows; h->splitGather = c.shape.splitGather; h->producerWarps = c.shape.producer;
	}

	// Compile and load the lattice-specialised kernel. Controlled by the environment: PFFRG_JIT=0 disables it,
	// PFFRG_JIT_MAX_TERMS (default 60000) bounds the straight-line code size (compile time grows with it).
	// With PFFRG_AUTOTUNE=1 small lattices (<= PFFRG_AUTOTUNE_MAX_TERMS terms, default 12000: a few seconds of compilation per shape) are AUTOTUNED:
	// up to five launch shapes (256-thread CTAs; 128-thread CTAs with the same RPA batch; 128-thread CTAs with two node groups,
	// four per SM; CTAs of four and of two 128-thread sub-CTAs, i.e. several work items sharing one RPA phase) are compiled and timed on a block of work items in the middle of the item range at a mid-mesh cutoff; the
	// fastest stays. Which shape wins depends on the lattice (measured: cubic-r7 the third, honeycomb-r7 the second, by 3-6 %).
	// Without it, or with an explicit shape override (PFFRG_JIT_NBT, PFFRG_THREADS), the first shape is used without timing.
	int setupJit(pffrg_context *h, const pffrg_desc *d, size_t smemMax, bool jitEnabled)
	{
		if (!jitEnabled) return PFFRG_OK;
		long maxTerms = 60000, tuneTerms = 12000;
		if (const char *e = getenv("PFFRG_JIT_MAX_TERMS")) maxTerms = atol(e);
		if (const char *e = getenv("PFFRG_AUTOTUNE_MAX_TERMS")) tuneTerms = atol(e);
		const auto t0 = std::chrono::steady_clock::now();
		// TRI: Gram form of the RPA phase in a run-time compiled kernel (rpaTriGram; PFFRG_RPA=table keeps the precompiled kernels with the
		// table-driven phase rpaTri8). A lattice it does not fit (more than 64 representatives, shared memory) stays on the precompiled kernels.
		if (h->core == TRI)
		{
			if (!wantTriGram(h->core) || h->nb != 8) return PFFRG_OK;
			JitShape shape = chooseTriGramShape(h->nw, h->L, h->groups, h->threads, smemMax);
			if (!shape.nb) return getenv("PFFRG_RPA") ? fail(PFFRG_ERR_UNSUPPORTED, "PFFRG_RPA=gram: the TRI Gram form does not fit this lattice (L %d)", h->L) : PFFRG_OK;
			JitCandidate c = { h->threads, h->groups, shape, nullptr, nullptr, 0.f };
			const int rc = compileCandidate(h, d, c);
			if (rc != PFFRG_OK) return rc;
			std::vector<unsigned short> blocks; std::vector<unsigned> terms; std::vector<int> seg;
			buildTriGramTables(d, h->L, shape.gramRows, h->threads / 32, blocks, terms, seg);
			CUDA_TRY(h->dTriBlocks.upload(blocks)); CUDA_TRY(h->dGramTerms.upload(terms)); CUDA_TRY(h->dGramSeg.upload(seg));
			h->triRounds = (int)(blocks.size() / shape.gramRows);
			h->triBlockCount = (int64_t)std::count_if(blocks.begin(), blocks.end(), [](unsigned short b) { return b != 0xffff; });
			h->gramWords = (int64_t)terms.size();
			adoptCandidate(h, c);
			h->jitCompileMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			return PFFRG_OK;
		}
		std::vector<JitCandidate> candidates;
		// Form of the RPA phase (SU2): PFFRG_RPA=gram -- Gram matrix over the staged nodes + one walk of the overlap list per phase
		// (rpaGram; no generated code, any lattice size); PFFRG_RPA=code -- lattice-specialised straight-line code. Default: the Gram form
		// for lattices with more than PFFRG_GRAM_MIN_TERMS (2000) merged overlap terms (as the warp-specialised kernel v4FlowBodySplit), where the straight-line code no longer fits the
		// instruction caches.
		if (h->core == SU2 || h->core == XYZ)
		{
			const char *form = getenv("PFFRG_RPA");
			if (wantGram(h->core, h->uniquePairs))
			{
				const GramLaunch launch = chooseGramLaunch(h->nw, h->L, h->Lp, h->groups, h->stride, h->threads, smemMax, h->uniquePairs, !form, h->core);
				const JitShape &shape = launch.shape;
				if (!shape.nb && form) return fail(PFFRG_ERR_UNSUPPORTED, "PFFRG_RPA=gram: no launch shape fits (threads %d, L %d)", h->threads, h->L);
				if (shape.nb)
				{
					JitCandidate c = { launch.threads, launch.groups, shape, nullptr, nullptr, 0.f };
					const int rc = compileCandidate(h, d, c);
					if (rc != PFFRG_OK) return rc;
					std::vector<unsigned> terms; std::vector<int> seg;
					buildGramTables(d, h->L, gramGeometry(h->core, h->L, h->Lp).lp, shape.gramRows, launch.reduceWarps, terms, seg);
					CUDA_TRY(h->dGramTerms.upload(terms)); CUDA_TRY(h->dGramSeg.upload(seg));
					h->gramWords = (int64_t)terms.size();
					adoptCandidate(h, c);
					h->jitCompileMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
					return PFFRG_OK;
				}
				// (no shape: the lattice stays on the straight-line code below)
			}
		}
		JitShape first = chooseJitShape(h->core, h->nw, h->L, h->groups, h->threads / 32, smemMax);
		if (!first.nb) return PFFRG_OK;
		int firstThreads = h->threads;
		if (const char *e = getenv("PFFRG_SUBCTAS")) // explicit shape: PFFRG_SUBCTAS sub-CTAs of PFFRG_THREADS threads, PFFRG_JIT_NBT / PFFRG_JIT_NB per sub-CTA
		{
			const int subs = atoi(e);
			if (subs > 1)
			{
				const JitShape fat = subCtaShape(h->core, h->nw, h->L, h->groups, h->threads / 32, subs, first.nbt, first.nb, getenv("PFFRG_JIT_MINBLOCKS") ? std::max(1, atoi(getenv("PFFRG_JIT_MINBLOCKS"))) : 1, smemMax);
				if (!fat.nb) return fail(PFFRG_ERR_UNSUPPORTED, "PFFRG_SUBCTAS=%d does not fit (shared memory / warps)", subs);
				first = fat; firstThreads = h->threads * subs;
			}
		}
		// CTAs per thread-block cluster (they rendezvous before every RPA phase, see clusterRendezvous): pairs by default for CTAs of one
		// work item -- measured on B200: pyrochlore-r8 243 -> 195 ms, cubic-r7 20.2 -> 18.0 ms, honeycomb-r7 XYZ 27.9 -> 25.2 ms with 2;
		// 4 and 8 are slower again (218 / 248 ms, 19.1 / 19.8 ms), CTAs of several sub-CTAs gain nothing
		int cluster = 2;
		if (const char *e = getenv("PFFRG_CLUSTER")) cluster = std::min(8, std::max(1, atoi(e)));
		first.cluster = (first.subs > 1 && !getenv("PFFRG_CLUSTER")) ? 1 : cluster;
		candidates.push_back({ firstThreads, h->groups, first, nullptr, nullptr, 0.f });
		const long terms = (long)buildRpaProgram(d, h->core, first.nbt, first.rpaWarps).terms.size();
		if (terms > maxTerms) return PFFRG_OK;
		// opt-in: the shapes differ in summation order (last-bit differences), so a run that must be reproducible bit for bit across
		// processes -- e.g. the sharded-vs-single-GPU comparison -- keeps the first shape
		bool tune = false;
		if (const char *e = getenv("PFFRG_AUTOTUNE")) tune = atoi(e) != 0 && terms <= tuneTerms && !getenv("PFFRG_JIT_NBT") && !getenv("PFFRG_THREADS") && !getenv("PFFRG_JIT_MINBLOCKS") && !getenv("PFFRG_SUBCTAS") && !getenv("PFFRG_CLUSTER");
		if (tune && h->threads > 128)
		{
			const int groups = std::max(1, 128 / h->stride), threads = std::max(64, (groups * h->stride + 31) / 32 * 32);
			if (threads < h->threads)
			{
				JitShape same = chooseJitShape(h->core, h->nw, h->L, groups, threads / 32, smemMax); same.cluster = cluster;
				if (same.nb) candidates.push_back({ threads, groups, same, nullptr, nullptr, 0.f });
				JitShape small = smallCtaShape(h->core, h->nw, h->L, groups, threads / 32, smemMax); small.cluster = cluster;
				if (small.nb) candidates.push_back({ threads, groups, small, nullptr, nullptr, 0.f });
				// several items per CTA (sub-CTAs of 128 threads, 32 staged nodes each, sharing one RPA phase): four in one CTA per SM,
				// two in two CTAs per SM (measured on B200: cubic-r7 20.2 -> 19.3 ms with 2 x 2, honeycomb-r7 XYZ 27.9 -> 26.2 ms with 4 x 1)
				for (int subs : { 4, 2 })
				{
					JitShape fat = subCtaShape(h->core, h->nw, h->L, groups, threads / 32, subs, 32, 16, 4 / subs, smemMax);
					if (!fat.nb && subs == 2) fat = subCtaShape(h->core, h->nw, h->L, groups, threads / 32, subs, 32, 16, 1, smemMax);
					if (fat.nb) candidates.push_back({ threads * subs, groups, fat, nullptr, nullptr, 0.f });
				}
			}
		}
		for (JitCandidate &c : candidates) { const int rc = compileCandidate(h, d, c); if (rc != PFFRG_OK) return rc; }
		size_t best = 0;
		if (candidates.size() > 1)
		{
			// timing run: zero vertex (timing is independent of the values), self energy zero, cutoff in the middle of the mesh
			const double cutoff = std::sqrt(h->mesh.front() * h->mesh.back());
			CUDA_TRY(cudaMemsetAsync(h->dV2.p, 0, h->nw * sizeof(double), h->stream));
			setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, cutoff);
			nodeTableKernel<<<h->nw, 128, sizeof(double) * (3 * h->nw + 2 * h->nodeStride), h->stream>>>(h->problem(), h->nodeTable(), h->dV2.p, h->dFlow2.p, h->dCutoff.p);
			CUDA_TRY(cudaGetLastError());
			const int64_t count = std::min<int64_t>(h->nf, 4736), begin = (h->nf - count) / 2; // 32 items per SM: whole waves for every candidate shape
			for (size_t k = 0; k < candidates.size(); ++k)
			{
				adoptCandidate(h, candidates[k]);
				for (int rep = 0; rep < 3; ++rep)
				{
					CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
					CUDA_TRY(launchFlowDispatch(h, begin, count));
					CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
					CUDA_TRY(cudaStreamSynchronize(h->stream));
					const float ms = elapsed(h->ev[0], h->ev[1]);
					if (rep == 1 || (rep == 2 && ms < candidates[k].ms)) candidates[k].ms = ms;
				}
				if (candidates[k].ms < candidates[best].ms) best = k;
			}
			CUDA_TRY(cudaMemsetAsync(h->dFlow4.p, 0, h->v4Elements() * sizeof(double), h->stream));
			CUDA_TRY(cudaMemsetAsync(h->dNan.p, 0, sizeof(int), h->stream));
			CUDA_TRY(cudaStreamSynchronize(h->stream));
			if (getenv("PFFRG_JIT_VERBOSE"))
				for (size_t k = 0; k < candidates.size(); ++k)
					fprintf(stderr, "[pffrg autotune] threads %d (%d sub-CTAs) nb %d nbt %d rpa warps %d ctas %d smem %zu: %.3f ms%s\n", candidates[k].threads, candidates[k].shape.subs, candidates[k].shape.nb, candidates[k].shape.nbt,
						candidates[k].shape.rpaWarps, candidates[k].shape.minBlocks, candidates[k].shape.smem, candidates[k].ms, k == best ? "  <- selected" : "");
		}
		for (size_t k = 0; k < candidates.size(); ++k) if (k != best && candidates[k].library) cudaLibraryUnload(candidates[k].library);
		adoptCandidate(h, candidates[best]);
		h->autotuned = (int)candidates.size();
		h->jitCompileMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		return PFFRG_OK;
	}

	// Device-internal order of the representative sites. A gather with the site/pair exchange flag reads, for site j, the entry of the
	// inverted site inv[j] (Lattice::getInvertedSites, src/SU2/SU2VertexTwoParticle.hpp:369-387): a permutation of the sites that moves
	// half of them on kagome / honeycomb / pyrochlore lattices and scatters a warp's 16-byte loads over up to 13 cache lines instead of 4
	// (ncu, pyrochlore-r8: 1.8 x the ideal number of L1 wavefronts over all gathers). Inversion is an involution, so the sites are
	// relabelled such that the two members of every pair {j, inv[j]} are neighbours (never straddling a 128-byte line): an exchanged
	// gather then touches the same lines as a plain one. Site 0 (the reference site: site-0 buffers, self-energy flow, correlations)
	// keeps index 0. All tables are relabelled once at creation; the reference order only exists at the boundary (import / export
	// kernels, bare couplings, correlation output). PFFRG_RELABEL=0 keeps the reference order.
	struct RelabelledDesc
	{
		std::vector<int> order, newOf; // order[new] = old, newOf[old] = new
		std::vector<int32_t> sitesRid, sitesPerm, invRid, invPerm, ovOff, ovRid1, ovRid2, ovPerm1, ovPerm2, rngFwd, rngInv;
		pffrg_desc view;
		bool identity = true;
		explicit RelabelledDesc(const pffrg_desc *d)
		{
			const int L = d->n_sites;
			order.resize(L); newOf.assign(L, -1);
			bool relabel = true;
			if (const char *e = getenv("PFFRG_RELABEL")) relabel = atoi(e) != 0;
			// involution check (anything else keeps the reference order)
			for (int j = 0; j < L && relabel; ++j) if (d->inverted_rid[d->inverted_rid[j]] != j) relabel = false;
			if (relabel && L > 0 && d->inverted_rid[0] != 0) relabel = false;
			if (relabel)
			{
				std::vector<int> singles, pairs;
				for (int j = 0; j < L; ++j)
				{
					const int k = d->inverted_rid[j];
					if (k == j) singles.push_back(j);
					else if (j < k) { pairs.push_back(j); pairs.push_back(k); }
				}
				// singles first (site 0 leads), an even number of them so that the pairs start on an even index; a surplus single goes last
				int n = 0, surplus = -1;
				if (singles.size() % 2 == 1 && !pairs.empty() && singles.size() > 1) { surplus = singles.back(); singles.pop_back(); }
				for (int j : singles) order[n++] = j;
				if (singles.size() % 2 == 1 && !pairs.empty()) { /* only site 0 is self-inverse: the pairs start at index 1 (one straddling pair per 8) */ }
				for (int j : pairs) order[n++] = j;
				if (surplus >= 0) order[n++] = surplus;
			}
			else for (int j = 0; j < L; ++j) order[j] = j;
			for (int n = 0; n < L; ++n) { newOf[order[n]] = n; if (order[n] != n) identity = false; }

			sitesRid.resize(L); invRid.resize(L); sitesPerm.resize(3 * L); invPerm.resize(3 * L);
			ovOff.assign(L + 1, 0);
			for (int n = 0; n < L; ++n)
			{
				const int j = order[n];
				sitesRid[n] = newOf[d->sites_rid[j]]; invRid[n] = newOf[d->inverted_rid[j]];
				for (int k = 0; k < 3; ++k) { sitesPerm[3 * n + k] = d->sites_perm[3 * j + k]; invPerm[3 * n + k] = d->inverted_perm[3 * j + k]; }
				ovOff[n + 1] = ovOff[n] + (d->overlap_offsets[j + 1] - d->overlap_offsets[j]);
			}
			const int total = ovOff[L];
			ovRid1.resize(total); ovRid2.resize(total); ovPerm1.resize(3 * (size_t)total); ovPerm2.resize(3 * (size_t)total);
			for (int n = 0; n < L; ++n)
			{
				const int j = order[n];
				for (int i = d->overlap_offsets[j], o = ovOff[n]; i < d->overlap_offsets[j + 1]; ++i, ++o)
				{
					ovRid1[o] = newOf[d->overlap_rid1[i]]; ovRid2[o] = newOf[d->overlap_rid2[i]];
					for (int k = 0; k < 3; ++k) { ovPerm1[3 * (size_t)o + k] = d->overlap_perm1[3 * (size_t)i + k]; ovPerm2[3 * (size_t)o + k] = d->overlap_perm2[3 * (size_t)i + k]; }
				}
			}
			rngFwd.resize(d->n_range); rngInv.resize(d->n_range);
			for (int k = 0; k < d->n_range; ++k) { rngFwd[k] = newOf[d->range_fwd_rid[k]]; rngInv[k] = newOf[d->range_inv_rid[k]]; }
			view = *d;
			view.sites_rid = sitesRid.data(); view.sites_perm = sitesPerm.data(); view.inverted_rid = invRid.data(); view.inverted_perm = invPerm.data();
			view.overlap_offsets = ovOff.data(); view.overlap_rid1 = ovRid1.data(); view.overlap_rid2 = ovRid2.data(); view.overlap_perm1 = ovPerm1.data(); view.overlap_perm2 = ovPerm2.data();
			view.range_fwd_rid = rngFwd.data(); view.range_inv_rid = rngInv.data();
		}
	};

	// (rid1, perm1, perm2, rid2) -> multiplicity of the overlap terms of one representative site (Lattice::getOverlap(rid),
	// src/Lattice.hpp:46-150); SU2 ignores the spin permutations
	std::map<std::tuple<int, int, int, int>, int> mergedOverlap(const pffrg_desc *d, int core, int rid)
	{
		std::map<std::tuple<int, int, int, int>, int> mult;
		for (int i = d->overlap_offsets[rid]; i < d->overlap_offsets[rid + 1]; ++i)
		{
			int p1 = core == SU2 ? 0 : packPerm(d->overlap_perm1 + 3 * i), p2 = core == SU2 ? 0 : packPerm(d->overlap_perm2 + 3 * i);
			mult[std::make_tuple(d->overlap_rid1[i], p1, p2, d->overlap_rid2[i])] += 1;
		}
		return mult;
	}

	// Term stream of the generic RPA phase (rpaGeneric): per representative site the merged overlap terms, sorted so that
	// equal (rid1, perm1, perm2) are adjacent; each such group starts with a header word, followed by one word per term.
	// Tasks (one per rid) are dealt to the RPA slots longest-first.
	// index of a packed spin permutation in the order of tri8Perm (pffrg_kernels.cuh)
	int permIndex(int packed)
	{
		for (int p = 0; p < 6; ++p)
			if (tri8Perm(p, 0) == (packed & 3) && tri8Perm(p, 1) == ((packed >> 2) & 3) && tri8Perm(p, 2) == ((packed >> 4) & 3)) return p;
		return 0;
	}

	void buildRpa(const pffrg_desc *d, int core, int nslots, int nbp, std::vector<unsigned> &words, std::vector<int4> &tasks, std::vector<int> &slotOff, int64_t &unique)
	{
		const int L = d->n_sites;
		std::vector<int4> perRid;
		unique = 0;
		for (int rid = 0; rid < L; ++rid)
		{
			const auto mult = mergedOverlap(d, core, rid);
			const int begin = (int)words.size();
			int lastR1 = -1, lastP1 = -1, lastP2 = -1;
			for (auto &kv : mult)
			{
				const int r1 = std::get<0>(kv.first), p1 = std::get<1>(kv.first), p2 = std::get<2>(kv.first), r2 = std::get<3>(kv.first);
				if (r1 != lastR1 || p1 != lastP1 || p2 != lastP2)
				{
					words.push_back(0x80000000u | (unsigned)(r1 * nbp) | ((unsigned)p1 << 16) | ((unsigned)p2 << 22));
					lastR1 = r1; lastP1 = p1; lastP2 = p2;
				}
				words.push_back((unsigned)(r2 * nbp) | ((unsigned)std::min(kv.second, 32767) << 16));
			}
			unique += (int64_t)mult.size();
			perRid.push_back(make_int4(rid, begin, (int)words.size(), 0));
		}
		std::vector<int> order(L);
		std::iota(order.begin(), order.end(), 0);
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return perRid[a].z - perRid[a].y > perRid[b].z - perRid[b].y; });
		std::vector<std::vector<int4>> bySlot(nslots);
		std::vector<long> load(nslots, 0);
		for (int r : order)
		{
			int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
			bySlot[best].push_back(perRid[r]);
			load[best] += perRid[r].z - perRid[r].y + 8;
		}
		slotOff.assign(1, 0);
		for (auto &sl : bySlot) { for (auto &t : sl) tasks.push_back(t); slotOff.push_back((int)tasks.size()); }
	}

	// Term stream of rpaTri8 (pffrg_kernels.cuh): per representative site the merged overlap terms sorted into runs of equal spin
	// permutations (p1, p2); inside a run, groups of equal rid1 (header word + one word per term). Each task = one rid:
	// {rid, position of its run descriptors, number of runs}; tasks are dealt to the warps longest-first.
	void buildRpaTri8(const pffrg_desc *d, int nWarps, std::vector<unsigned> &words, std::vector<int4> &tasks, std::vector<int> &slotOff, int64_t &unique)
	{
		const int L = d->n_sites;
		struct Entry { int p1, p2, r1, r2, mult; };
		std::vector<int4> perRid; std::vector<long> cost;
		unique = 0;
		for (int rid = 0; rid < L; ++rid)
		{
			std::vector<Entry> entries;
			for (auto &kv : mergedOverlap(d, TRI, rid))
				entries.push_back({ permIndex(std::get<1>(kv.first)), permIndex(std::get<2>(kv.first)), std::get<0>(kv.first), std::get<3>(kv.first), kv.second });
			unique += (int64_t)entries.size();
			std::stable_sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return std::make_tuple(a.p1, a.p2, a.r1, a.r2) < std::make_tuple(b.p1, b.p2, b.r1, b.r2); });
			std::vector<unsigned> descriptors, body;
			std::vector<size_t> descriptorBodyOffset;
			size_t i = 0;
			while (i < entries.size())
			{
				size_t j = i;
				while (j < entries.size() && entries[j].p1 == entries[i].p1 && entries[j].p2 == entries[i].p2) ++j;
				unsigned groups = 0;
				while (body.size() & 3) body.push_back(0u); // records are read in pairs as 16-byte words
				descriptorBodyOffset.push_back(body.size());
				std::vector<unsigned> extras;
				for (size_t k = i; k < j;)
				{
					size_t m = k;
					while (m < j && entries[m].r1 == entries[k].r1 && m - k < 1024) ++m;
					body.push_back((unsigned)(entries[k].r1 * TRI8_RID_STRIDE) | ((unsigned)(m - k - 1) << 22));
					for (size_t t = k; t < m; ++t)
						(t == k ? body : extras).push_back((unsigned)(entries[t].r2 * TRI8_RID_STRIDE) | ((unsigned)std::min(entries[t].mult, 32767) << 16));
					++groups; k = m;
				}
				if (groups & 1) { body.push_back(0u); body.push_back(0u); } // odd run: a record with multiplicity 0 completes the last pair
				body.insert(body.end(), extras.begin(), extras.end());
				descriptors.push_back((unsigned)entries[i].p1 | ((unsigned)entries[i].p2 << 3) | (groups << 6));
				i = j;
			}
			const int first = (int)words.size();
			const size_t bodyBase = (words.size() + 2 * descriptors.size() + 3) / 4 * 4;
			for (size_t q = 0; q < descriptors.size(); ++q) { words.push_back(descriptors[q]); words.push_back((unsigned)(bodyBase + descriptorBodyOffset[q])); }
			words.resize(bodyBase, 0u);
			words.insert(words.end(), body.begin(), body.end());
			perRid.push_back(make_int4(rid, first, (int)descriptors.size(), 0));
			cost.push_back((long)body.size() + 16 * (long)descriptors.size() + 8);
		}
		std::vector<int> order(L);
		std::iota(order.begin(), order.end(), 0);
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
		std::vector<std::vector<int4>> byWarp(nWarps);
		std::vector<long> load(nWarps, 0);
		for (int r : order)
		{
			const int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
			byWarp[best].push_back(perRid[r]);
			load[best] += cost[r];
		}
		slotOff.assign(1, 0);
		for (auto &w : byWarp) { for (auto &t : w) tasks.push_back(t); slotOff.push_back((int)tasks.size()); }
	}

	// the RPA sum as a list of multiply-adds over staged operands [channel][rid], for the code generator
	RpaProgram buildRpaProgram(const pffrg_desc *d, int core, int nb /* nodes per RPA phase */, int warps /* RPA warps */)
	{
		const int L = d->n_sites, C = channelsOf(core);
		RpaProgram p;
		p.nb = nb; p.warps = warps;
		p.operandStride = nb + 1;
		p.operandBOffset = (long)C * L * (nb + 1);
		p.outputCopyStride = (long)C * L;
		if (core == SU2)
		{
			// both channels share the term structure: run them as two lane groups of one warp
			p.variants = 2; p.lanesPerVariant = 16; p.nOutputs = L;
			p.variantOperandStride = (long)L * (nb + 1); p.variantOutputStride = L;
			for (int rid = 0; rid < L; ++rid)
				for (auto &kv : mergedOverlap(d, core, rid)) p.terms.push_back({ rid, std::get<0>(kv.first), std::get<3>(kv.first), kv.second });
		}
		else
		{
			p.variants = 1; p.lanesPerVariant = std::min(nb, 32); p.nOutputs = C * L;
			for (int rid = 0; rid < L; ++rid)
				for (auto &kv : mergedOverlap(d, core, rid))
				{
					const int r1 = std::get<0>(kv.first), p1 = std::get<1>(kv.first), p2 = std::get<2>(kv.first), r2 = std::get<3>(kv.first);
					for (int c = 0; c < 3; ++c) p.terms.push_back({ c * L + rid, ((p1 >> (2 * c)) & 3) * L + r1, ((p2 >> (2 * c)) & 3) * L + r2, kv.second });
					p.terms.push_back({ 3 * L + rid, 3 * L + r1, 3 * L + r2, kv.second });
				}
		}
		return p;
	}

	// Word stream of one reduction pass (gramReduce / triGramReduce): `lists[k]` = the words of output key k (ascending keys = ascending lane
	// index inside a chunk). Appended to `terms`:
	// - every list is padded to a multiple of 8 words with padWord(key, class) (multiplicity 0), so that the 8 consecutive words a lane reads
	//   belong to one key;
	// - the keys are dealt to the `warps` warps as contiguous ranges of about equal length; every range starts on a multiple of 256 words and
	//   is padded to whole chunks of 256 words: ranges[2 w] = {begin, end};
	// - inside a list the words are ordered so that in every step k the lanes that are served together (a quarter warp for 16-byte entries,
	//   classBits = 3; a half warp for 8-byte entries, classBits = 4; words 8 l + k of the group) address different bank groups of the
	//   Gram block (offset mod 2^classBits) where the list allows it. Returns (sets, sum of conflict degrees) for the statistics.
	template <typename PadWord>
	std::pair<double, double> layoutReductionPass(std::vector<std::vector<unsigned>> &lists, int classBits, unsigned offsetMask, int warps, PadWord padWord,
	                                                std::vector<unsigned> &terms, int *ranges)
	{
		const int classes = 1 << classBits, group = 8 * classes; // words served together per step: `classes` lanes x 8 words
		const int nKeys = (int)lists.size();
		std::vector<int> padded(nKeys, 0);
		long total = 0;
		for (int k = 0; k < nKeys; ++k) { padded[k] = ((int)lists[k].size() + 7) / 8 * 8; total += padded[k]; }
		double sets = 0.0, degree = 0.0;
		int key = 0; long done = 0;
		for (int w = 0; w < warps; ++w)
		{
			while (terms.size() % 256) terms.push_back(0u);
			const size_t begin = terms.size();
			ranges[2 * w] = (int)begin;
			const long target = (w + 1 == warps) ? total : total * (w + 1) / warps;
			int lastKey = key < nKeys ? key : nKeys - 1;
			while (key < nKeys && (done < target || padded[key] == 0))
			{
				std::vector<std::vector<unsigned>> byClass(classes);
				for (unsigned word : lists[key]) byClass[(word & offsetMask) & (classes - 1)].push_back(word);
				std::map<size_t, unsigned> used; // set id -> classes taken by this list
				for (int i = 0; i < padded[key]; ++i)
				{
					const size_t pos = terms.size() - begin;
					unsigned &mask = used[(pos / group) * 8 + pos % 8];
					int best = -1;
					for (int c = 0; c < classes; ++c)
						if (!byClass[c].empty() && !(mask & (1u << c)) && (best < 0 || byClass[c].size() > byClass[best].size())) best = c;
					if (best < 0) for (int c = 0; c < classes; ++c) if (!byClass[c].empty() && (best < 0 || byClass[c].size() > byClass[best].size())) best = c;
					if (best >= 0) { terms.push_back(byClass[best].back()); byClass[best].pop_back(); mask |= 1u << best; }
					else
					{
						int c = 0; while (c < classes - 1 && (mask & (1u << c))) ++c;
						terms.push_back(padWord(key, c));
						mask |= 1u << c;
					}
				}
				done += padded[key]; lastKey = key; ++key;
			}
			while ((terms.size() - begin) % 256) terms.push_back(padWord(lastKey < 0 ? 0 : lastKey, 0));
			ranges[2 * w + 1] = (int)terms.size();
			for (size_t base = begin; base < terms.size(); base += group)
				for (int k = 0; k < 8; ++k)
				{
					std::vector<int> count(classes, 0); int occupied = 0;
					for (int l = 0; l < classes; ++l) ++count[(terms[base + 8 * l + k] & offsetMask) & (classes - 1)];
					for (int c = 0; c < classes; ++c) occupied += count[c] > 0;
					sets += 1.0; degree += (double)classes / occupied;
				}
		}
		return { sets, degree };
	}

	// Term tables of gramReduce (SU2), "lane-serial" layout. For every block of PB rows of the Gram matrix the merged overlap terms
	// (rid1, rid2, multiplicity) of all representative sites rid with rid1 in the block become 32-bit words
	//     (rid1 - block * PB) * (Lp + 1) + rid2  |  rid << 14  |  multiplicity << 22  |  flush << 31.
	// Whole rid lists are dealt to the warps (longest first, to the least loaded warp: single writer per output and block). Inside a warp
	// every rid list is padded to whole groups of 4 words and the groups of all its lists are concatenated; lane l walks the groups
	// [l T4, (l + 1) T4) of that sequence SERIALLY, accumulating in registers; the last group of a rid carries the flush flag (the lane adds
	// its sum to the output and starts over), and what a lane holds at the end of its range (a rid that continues in the next lane) is
	// combined by one segmented scan per (block, warp). No cross-lane traffic per term, no padding beyond the groups of 4: the previous
	// layout (8 words per lane and chunk + a 5-level scan per chunk) cost 27 instructions per term and lane, this one ~8.
	// Stored as [group][lane][4 words] (one coalesced 16-byte load per lane and group); seg[2 (block * warps + w)] = {first word, T4}.
	// Inside a lane's piece of a rid list the words are ordered so that the 8 lanes of a quarter warp address different 16-byte bank
	// groups of the Gram block (offset mod 8) in every step where the lists allow it.
	// conflictDegree (optional): average number of words per occupied bank group over all (quarter warp, step) sets; 1 = conflict free.
	void buildGramTables(const pffrg_desc *d, int L, int Lp, int PB, int warps, std::vector<unsigned> &terms, std::vector<int> &seg, double *conflictDegree)
	{
		const int blocks = (Lp + PB - 1) / PB, maxMult = 511;
		const int core = d->core == XYZ ? XYZ : SU2;
		// XYZ: operand index = channel * L + site, output = channel * L + rid for the spin channels (read from G.x), 3 L + rid for the density
		// channel (from G.y, which is non-zero in the corner of the two c = 0 operand ranges only); what the other component of an entry adds
		// to the other half of the outputs is never read (v4FlowBodySplit's epilogue)
		const int outputs = core == XYZ ? 4 * L : L;
		std::vector<std::map<std::tuple<int, int, int, int>, int>> merged(outputs); // per output: (operand p, 0, 0, operand q) -> multiplicity
		for (int rid = 0; rid < L; ++rid)
		{
			if (core == SU2) { merged[rid] = mergedOverlap(d, SU2, rid); continue; }
			for (auto &kv : mergedOverlap(d, XYZ, rid))
			{
				const int r1 = std::get<0>(kv.first), p1 = std::get<1>(kv.first), p2 = std::get<2>(kv.first), r2 = std::get<3>(kv.first);
				for (int c = 0; c < 3; ++c) merged[c * L + rid][std::make_tuple(((p1 >> (2 * c)) & 3) * L + r1, 0, 0, ((p2 >> (2 * c)) & 3) * L + r2)] += kv.second;
				merged[3 * L + rid][std::make_tuple(r1, 0, 0, r2)] += kv.second;
			}
		}
		seg.assign((size_t)2 * blocks * warps, 0);
		terms.clear();
		double sets = 0.0, degree = 0.0;
		for (int blk = 0; blk < blocks; ++blk)
		{
			std::vector<std::vector<unsigned>> lists(outputs);
			for (int rid = 0; rid < outputs; ++rid)
				for (auto &kv : merged[rid])
				{
					const int p = std::get<0>(kv.first), q = std::get<3>(kv.first);
					if (p / PB != blk) continue;
					const unsigned offset = (unsigned)((p - blk * PB) * (Lp + 1) + q); // row stride of the Gram block: gramcfg::LpG
					for (int m = kv.second; m > 0; m -= maxMult) lists[rid].push_back(offset | ((unsigned)rid << 14) | ((unsigned)std::min(m, maxMult) << 22));
				}
			// whole lists to warps: longest first, to the least loaded warp
			std::vector<int> order;
			for (int rid = 0; rid < outputs; ++rid) if (!lists[rid].empty()) order.push_back(rid);
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lists[a].size() > lists[b].size(); });
			std::vector<std::vector<int>> owned(warps);
			std::vector<long> load(warps, 0);
			for (int rid : order)
			{
				const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
				owned[w].push_back(rid); load[w] += ((long)lists[rid].size() + 3) / 4;
			}
			for (int w = 0; w < warps; ++w)
			{
				std::sort(owned[w].begin(), owned[w].end());
				// the sequence of groups: (rid, last group of its list?)
				std::vector<std::pair<int, bool>> groups;
				for (int rid : owned[w])
				{
					const int n = ((int)lists[rid].size() + 3) / 4;
					for (int g = 0; g < n; ++g) groups.push_back({ rid, g + 1 == n });
				}
				const int T4 = ((int)groups.size() + 31) / 32;
				const size_t begin = terms.size();
				seg[2 * ((size_t)blk * warps + w)] = (int)begin; seg[2 * ((size_t)blk * warps + w) + 1] = T4;
				terms.resize(begin + (size_t)T4 * 128, 0u);
				const int padRid = owned[w].empty() ? 0 : owned[w].back();
				// remaining words of every list by bank class
				std::map<int, std::vector<std::vector<unsigned>>> pool;
				for (int rid : owned[w]) { auto &byClass = pool[rid]; byClass.resize(8); for (unsigned word : lists[rid]) byClass[word & 7u].push_back(word); }
				std::vector<unsigned char> used((size_t)4 * T4 * 4, 0); // [quarter][group][k] -> classes taken
				for (int lane = 0; lane < 32; ++lane)
					for (int g = 0; g < T4; ++g)
					{
						const size_t gi = (size_t)lane * T4 + g;
						const bool real = gi < groups.size();
						const int rid = real ? groups[gi].first : padRid;
						for (int k = 0; k < 4; ++k)
						{
							unsigned char &mask = used[((size_t)(lane / 8) * T4 + g) * 4 + k];
							unsigned word;
							int best = -1;
							if (real)
							{
								auto &byClass = pool[rid];
								for (int c = 0; c < 8; ++c) if (!byClass[c].empty() && !(mask & (1u << c)) && (best < 0 || byClass[c].size() > byClass[best].size())) best = c;
								if (best < 0) for (int c = 0; c < 8; ++c) if (!byClass[c].empty() && (best < 0 || byClass[c].size() > byClass[best].size())) best = c;
								if (best >= 0) { word = byClass[best].back(); byClass[best].pop_back(); }
							}
							if (best < 0)
							{
								// padding word: multiplicity 0, a free bank class (offsets 0..7 exist: a block has at least 8 rows)
								int c = 0; while (c < 7 && (mask & (1u << c))) ++c;
								word = (unsigned)c | ((unsigned)rid << 14); best = c;
							}
							mask |= (unsigned char)(1u << best);
							if (real && groups[gi].second && k == 3) word |= 1u << 31;
							terms[begin + ((size_t)g * 32 + lane) * 4 + k] = word;
						}
					}
				for (auto &kv : pool) for (auto &v : kv.second) if (!v.empty()) abort(); // every word placed (a list's groups hold all of its words)
				for (int q = 0; q < 4; ++q)
					for (int g = 0; g < T4; ++g)
						for (int k = 0; k < 4; ++k)
						{
							int count[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, occupied = 0;
							for (int l = 8 * q; l < 8 * q + 8; ++l) ++count[terms[begin + ((size_t)g * 32 + l) * 4 + k] & 7u];
							for (int c = 0; c < 8; ++c) occupied += count[c] > 0;
							sets += 1.0; degree += 8.0 / occupied;
						}
			}
		}
		if (conflictDegree) *conflictDegree = sets > 0 ? degree / sets : 1.0;
	}

	// Tables of the Gram form of the TRI RPA phase (rpaTriGram / triGramReduce, pffrg_kernels.cuh). The needed channel pairs (c1, c2) =
	// ((p1 mu, p1 k), (p2 k, p2 nu)) over all overlap terms and (mu, k, nu) are worked off in rounds of `resident` blocks (most terms first):
	// blocks[round * resident + slot] = c1 | c2 << 4 (0xffff: none). Per round one pass of words
	//     slot * GBLK + rid1 * GS + rid2  |  out << 13  |  minus << 23  |  multiplicity << 24,      out = (4 mu + nu) * L + rid,
	// with eta(mu,k,nu) = tri::rpa(mu,k,nu).sign (src/TRI/TRIFrgCore.cpp:739-1252; the factor 2 and the node weight are folded into the
	// staged operand A) and terms of equal (block, rid1, rid2, out) merged.
	void buildTriGramTables(const pffrg_desc *d, int L, int resident, int warps, std::vector<unsigned short> &blocks, std::vector<unsigned> &terms, std::vector<int> &seg, double *conflictDegree)
	{
		const int LT = (L + 7) / 8, LpT = 8 * LT, GS = LpT + 1, GBLK = LpT * GS;
		// (c1 | c2 << 4) -> (rid1, rid2, out) -> signed multiplicity
		std::map<int, std::map<std::tuple<int, int, int>, int>> byBlock;
		for (int rid = 0; rid < L; ++rid)
			for (auto &kv : mergedOverlap(d, TRI, rid))
			{
				const int r1 = std::get<0>(kv.first), p1 = std::get<1>(kv.first), p2 = std::get<2>(kv.first), r2 = std::get<3>(kv.first);
				auto perm = [](int p, int i) { return i == 3 ? 3 : (p >> (2 * i)) & 3; };
				for (int mu = 0; mu < 4; ++mu) for (int k = 0; k < 4; ++k) for (int nu = 0; nu < 4; ++nu)
				{
					const int c1 = 4 * perm(p1, mu) + perm(p1, k), c2 = 4 * perm(p2, k) + perm(p2, nu);
					byBlock[c1 | (c2 << 4)][std::make_tuple(r1, r2, (4 * mu + nu) * L + rid)] += (int)tri::rpa(mu, k, nu).sign * kv.second;
				}
			}
		std::vector<int> order;
		for (auto &kv : byBlock) order.push_back(kv.first);
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return byBlock[a].size() > byBlock[b].size(); });
		const int rounds = ((int)order.size() + resident - 1) / resident;
		blocks.assign((size_t)rounds * resident, (unsigned short)0xffff);
		seg.assign((size_t)2 * rounds * warps, 0);
		terms.clear();
		double sets = 0.0, degree = 0.0;
		for (int round = 0; round < rounds; ++round)
		{
			std::vector<std::vector<unsigned>> lists((size_t)16 * L);
			for (int slot = 0; slot < resident && round * resident + slot < (int)order.size(); ++slot)
			{
				const int pair = order[round * resident + slot];
				blocks[(size_t)round * resident + slot] = (unsigned short)pair;
				for (auto &kv : byBlock[pair])
				{
					const int r1 = std::get<0>(kv.first), r2 = std::get<1>(kv.first), out = std::get<2>(kv.first);
					const unsigned offset = (unsigned)(slot * GBLK + r1 * GS + r2);
					for (int m = std::abs(kv.second); m > 0; m -= 255)
						lists[out].push_back(offset | ((unsigned)out << 13) | ((kv.second < 0 ? 1u : 0u) << 23) | ((unsigned)std::min(m, 255) << 24));
				}
			}
			const auto stats = layoutReductionPass(lists, 4, (1u << 13) - 1u, warps, [](int out, int c) { return (unsigned)c | ((unsigned)out << 13); }, terms, seg.data() + (size_t)2 * round * warps);
			sets += stats.first; degree += stats.second;
		}
		if (conflictDegree) *conflictDegree = sets > 0 ? degree / sets : 1.0;
	}

	// node counts per mesh frequency at the current cutoff (host copy of what nodeTableKernel enumerates)
	std::vector<int> hostNodeCounts(const pffrg_context *h)
	{
		std::vector<int> c(h->nw);
		for (int i = 0; i < h->nw; ++i) c[i] = nodeCount(h->mesh.data(), h->nw, h->cutoff, h->mesh[i]);
		return c;
	}

	// Contiguous cost-balanced item ranges (replaces the dynamic master/worker chunking of src/lib/LoadManager.hpp:796-852).
	// Items are ordered su-major, t-minor; the cost of item (so, uo, t) is n(so) + n(uo) ladder evaluations and n(t)
	// t-channel evaluations, n(x) the quadrature node count of transfer frequency x at the current cutoff.
	// Modelled cost of the items [0, item): items are ordered su-major, t-minor; item (so, uo, t) costs n(so) + n(uo) ladder
	// evaluations and n(t) t-channel evaluations, n(x) the quadrature node count of transfer frequency x at the current cutoff.
	struct CostModel
	{
		int nw; double costSU, costT;
		std::vector<double> prefix;  // cost of the su blocks [0, b)
		std::vector<double> prefixT; // sum_{t' < t} n(t')
		const std::vector<int> *counts;
		CostModel(int core, int nw_, double L, double rpaTerms, const std::vector<int> &c) : nw(nw_), counts(&c)
		{
			const CoreModel m = modelOf(core);
			costSU = (16.0 * m.C + m.ladderTerms) * L;
			costT = (16.0 * m.C + m.localTerms) * L + 2.0 * m.rpaTerms * rpaTerms;
			prefixT.assign(nw + 1, 0.0);
			for (int t = 0; t < nw; ++t) prefixT[t + 1] = prefixT[t] + c[t];
			prefix.assign((size_t)nw * (nw + 1) / 2 + 1, 0.0);
			int64_t su = 0;
			for (int so = 0; so < nw; ++so)
				for (int uo = 0; uo <= so; ++uo, ++su)
					prefix[su + 1] = prefix[su] + nw * (double)(c[so] + c[uo]) * costSU + prefixT[nw] * costT;
		}
		int64_t items() const { return (int64_t)(prefix.size() - 1) * nw; }
		double upTo(int64_t item) const
		{
			const int64_t blk = item / nw; const int t = (int)(item - blk * nw);
			if (t == 0) return prefix[blk];
			int so = (int)((std::sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
			while ((int64_t)(so + 1) * (so + 2) / 2 <= blk) ++so;
			while ((int64_t)so * (so + 1) / 2 > blk) --so;
			const int uo = (int)(blk - (int64_t)so * (so + 1) / 2);
			return prefix[blk] + t * (double)((*counts)[so] + (*counts)[uo]) * costSU + prefixT[t] * costT;
		}
	};

	// Contiguous cost-balanced item ranges (replaces the dynamic master/worker chunking of src/lib/LoadManager.hpp:796-852).
	// Without feedback the modelled cost is split evenly. With feedback -- the item boundaries and the measured flow-kernel times of
	// the previous step, identical on all ranks -- the modelled cost is weighted by the measured time per modelled unit of the
	// previous rank interval an item lies in (a piecewise constant density): what the model does not know (cache locality of the
	// gathers varies along the item axis; measured on 4 GPUs: 52 ms on the first rank against 63 ms on the slowest at pyrochlore-r8
	// with equal modelled cost) is corrected from one step to the next, the way the reference's LoadManager adapts its chunk sizes
	// to the measured throughput of its workers.
	std::vector<int64_t> planPartition(int core, int nw, double L, double rpaTerms, const std::vector<int> &counts, int nRanks,
	                                   const std::vector<int64_t> *prevBounds = nullptr, const std::vector<double> *prevTimes = nullptr)
	{
		const CostModel model(core, nw, L, rpaTerms, counts);
		const int64_t nf = model.items();
		// density pieces: [edge[k], edge[k+1]) with weight rho[k]
		std::vector<int64_t> edge = { 0, nf };
		std::vector<double> rho = { 1.0 };
		if (prevBounds && prevTimes && (int)prevBounds->size() == nRanks + 1 && (int)prevTimes->size() == nRanks && nRanks > 1)
		{
			std::vector<double> r(nRanks, 0.0); double mean = 0.0; int valid = 0;
			for (int k = 0; k < nRanks; ++k)
			{
				const double c = model.upTo((*prevBounds)[k + 1]) - model.upTo((*prevBounds)[k]);
				if (c > 0.0 && (*prevTimes)[k] > 0.0) { r[k] = (*prevTimes)[k] / c; mean += r[k]; ++valid; }
			}
			if (valid == nRanks)
			{
				mean /= nRanks;
				for (double &x : r) x = std::min(2.0, std::max(0.5, x / mean)); // bounded correction
				edge = *prevBounds; edge.front() = 0; edge.back() = nf; rho = r;
			}
		}
		auto weightedUpTo = [&](int64_t item)
		{
			double w = 0.0;
			for (size_t k = 0; k + 1 < edge.size(); ++k)
			{
				if (item <= edge[k]) break;
				w += rho[k] * (model.upTo(std::min(item, edge[k + 1])) - model.upTo(edge[k]));
			}
			return w;
		};
		const double total = weightedUpTo(nf);
		std::vector<int64_t> bounds(nRanks + 1, 0);
		bounds[nRanks] = nf;
		for (int r = 1; r < nRanks; ++r)
		{
			const double target = total * r / nRanks;
			int64_t lo = bounds[r - 1], hi = nf; // smallest item count whose weighted cost reaches the target
			while (lo < hi)
			{
				const int64_t mid = (lo + hi) / 2;
				if (weightedUpTo(mid) < target) lo = mid + 1; else hi = mid;
			}
			bounds[r] = lo;
		}
		return bounds;
	}

	void partitionItems(pffrg_context *h, const std::vector<int> &counts)
	{
		// feedback from the previous step's kernel times (multi-GPU runs; PFFRG_BALANCE=0 keeps the static split)
		const bool feedback = h->nRanks > 1 && h->balance && (int)h->rankTimes.size() == h->nRanks && (int)h->bounds.size() == h->nRanks + 1;
		const std::vector<int64_t> prev = h->bounds;
		// RPA work per t-channel node in units of the cost model (overlap terms of the indexed forms)
		double rpaTerms = (double)h->uniquePairs;
		if (h->gramRows > 0 && h->core == SU2) rpaTerms = (double)h->L * h->L;
		if (h->gramRows > 0 && h->core == XYZ) rpaTerms = 9.0 * h->L * h->L / 2.0; // (3 L)^2 double2 entries per node against 4 channels per term of the indexed form
		if (h->gramRows > 0 && h->core == TRI) { const double LpT = (double)((h->L + 7) / 8 * 8); rpaTerms = (double)h->triBlockCount * LpT * LpT * 2.0 / 128.0; }
		h->bounds = planPartition(h->core, h->nw, h->L, rpaTerms, counts, h->nRanks, feedback ? &prev : nullptr, feedback ? &h->rankTimes : nullptr);
	}

	void fillStats(pffrg_context *h, const std::vector<int> &counts, int64_t begin, int64_t end)
	{
		const CoreModel m = modelOf(h->core);
		const int nw = h->nw;
		int64_t sumT = 0; for (int t = 0; t < nw; ++t) sumT += counts[t];
		std::vector<int64_t> prefixT(nw + 1, 0); for (int t = 0; t < nw; ++t) prefixT[t + 1] = prefixT[t] + counts[t];
		int64_t nS = 0, nT = 0, nU = 0;
		// items are ordered su-major, t-minor: whole (s,u) blocks contribute nw * (n(so) + n(uo)) ladder and sum_t n(t) t-channel evaluations
		int64_t su = 0;
		for (int so = 0; so < nw; ++so)
			for (int uo = 0; uo <= so; ++uo, ++su)
			{
				const int64_t lo = std::max<int64_t>(begin, su * nw), hi = std::min<int64_t>(end, (su + 1) * nw);
				if (hi <= lo) continue;
				nS += (hi - lo) * counts[so]; nU += (hi - lo) * counts[uo];
				nT += prefixT[hi - su * nw] - prefixT[lo - su * nw];
			}
		const double L = h->L, C = m.C;
		h->stats.items = end - begin;
		h->stats.kernel_evals = nS + nT + nU;
		h->stats.kernel_evals_t = nT;
		h->stats.alg_bytes = (double)(nS + nT + nU) * 128.0 * L * C + (double)nT * 128.0 * C + (double)(end - begin) * L * C * 24.0;
		const double fmaSU = 16.0 * L * C + m.ladderTerms * L;
		const double fmaT = 16.0 * L * C + (double)m.rpaTerms * h->overlapTotal + m.localTerms * L;
		h->stats.alg_flops = 2.0 * ((double)(nS + nU) * fmaSU + (double)nT * fmaT);
		// as executed: merged overlap terms; Gram form: an L x L update per node and one walk of the merged terms per RPA phase
		double rpaFma = (double)nT * m.rpaTerms * (double)h->uniquePairs;
		if (h->gramRows > 0)
		{
			std::vector<int64_t> prefixR(nw + 1, 0); // RPA phases of the t values below an index
			for (int t = 0; t < nw; ++t) prefixR[t + 1] = prefixR[t] + (counts[t] + h->nbt - 1) / h->nbt;
			int64_t phases = 0, blk = 0;
			for (int so = 0; so < nw; ++so)
				for (int uo = 0; uo <= so; ++uo, ++blk)
				{
					const int64_t lo = std::max<int64_t>(begin, blk * nw), hi = std::min<int64_t>(end, (blk + 1) * nw);
					if (hi > lo) phases += prefixR[hi - blk * nw] - prefixR[lo - blk * nw];
				}
			if (h->core == TRI)
			{
				const double LpT = (double)((h->L + 7) / 8 * 8);
				rpaFma = (double)nT * 2.0 * (double)h->triBlockCount * LpT * LpT + (double)phases * (double)h->gramWords;
			}
			else rpaFma = (double)nT * C * L * L + (double)phases * C * (double)h->uniquePairs;
		}
		h->stats.exec_flops = 2.0 * ((double)(nS + nU) * fmaSU + (double)nT * (16.0 * L * C + m.localTerms * L) + rpaFma);
	}

	int exchangeSlices(pffrg_context *h, double *buffer)
	{
		if (h->nRanks <= 1) return PFFRG_OK;
		NCCL_TRY(nccl().GroupStart());
		for (int r = 0; r < h->nRanks; ++r)
		{
			size_t off = (size_t)h->bounds[r] * h->RL, cnt = (size_t)(h->bounds[r + 1] - h->bounds[r]) * h->RL;
			if (cnt == 0) continue;
			NCCL_TRY(nccl().Broadcast(buffer + off, buffer + off, cnt, ncclDouble, r, h->comm, h->stream));
		}
		NCCL_TRY(nccl().GroupEnd());
		return PFFRG_OK;
	}

	// rows [rowBegin, rowBegin + rows) of the vertex between the host arrays (reference layout) and a device buffer (device layout)
	template <typename T>
	int importArrays(pffrg_context *h, const void *const *src, double *dst, int64_t rowBegin, int64_t rows)
	{
		const size_t per = (size_t)h->L * (h->core == TRI ? 16 : 1), len = (size_t)rows * per;
		if (rows <= 0) return PFFRG_OK;
		if (h->dStaging.n * sizeof(double) < len * sizeof(T)) CUDA_TRY(h->dStaging.alloc((len * sizeof(T) + 7) / 8));
		T *staging = reinterpret_cast<T *>(h->dStaging.p);
		for (int a = 0; a < h->nArrays; ++a)
		{
			if (!src[a]) return fail(PFFRG_ERR_ARGUMENT, "vertex array %d is null", a);
			CUDA_TRY(cudaMemcpyAsync(staging, static_cast<const T *>(src[a]) + (size_t)rowBegin * per, len * sizeof(T), cudaMemcpyHostToDevice, h->stream));
			importKernel<T><<<1184, 256, 0, h->stream>>>(staging, dst, (size_t)rowBegin, (size_t)rows, h->L, h->Lp, h->RL, vectorWidth(h->core), h->core == TRI ? 0 : a, h->core == TRI ? 16 : 1, h->dSiteNewOf.p);
			CUDA_TRY(cudaGetLastError());
		}
		return PFFRG_OK;
	}

	template <typename T>
	int exportArrays(pffrg_context *h, const double *src, void *const *dst, int64_t rowBegin, int64_t rows)
	{
		const size_t per = (size_t)h->L * (h->core == TRI ? 16 : 1), len = (size_t)rows * per;
		if (rows <= 0) return PFFRG_OK;
		if (h->dStaging.n * sizeof(double) < len * sizeof(T)) CUDA_TRY(h->dStaging.alloc((len * sizeof(T) + 7) / 8));
		T *staging = reinterpret_cast<T *>(h->dStaging.p);
		for (int a = 0; a < h->nArrays; ++a)
		{
			if (!dst[a]) continue;
			exportKernel<T><<<1184, 256, 0, h->stream>>>(src, staging, (size_t)rowBegin, (size_t)rows, h->L, h->Lp, h->RL, vectorWidth(h->core), h->core == TRI ? 0 : a, h->core == TRI ? 16 : 1, h->dSiteNewOf.p);
			CUDA_TRY(cudaGetLastError());
			CUDA_TRY(cudaMemcpyAsync(static_cast<T *>(dst[a]) + (size_t)rowBegin * per, staging, len * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
		}
		CUDA_TRY(cudaStreamSynchronize(h->stream));
		return PFFRG_OK;
	}

	template <typename T>
	int importVector(pffrg_context *h, const void *src, double *dst, int n)
	{
		if (h->dVecStaging.n < (size_t)n) CUDA_TRY(h->dVecStaging.alloc(n));
		T *staging = reinterpret_cast<T *>(h->dVecStaging.p);
		CUDA_TRY(cudaMemcpyAsync(staging, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
		convertKernel<T><<<1, 128, 0, h->stream>>>(staging, dst, n);
		CUDA_TRY(cudaGetLastError());
		return PFFRG_OK;
	}
	template <typename T>
	int exportVector(pffrg_context *h, const double *src, void *dst, int n)
	{
		if (h->dVecStaging.n < (size_t)n) CUDA_TRY(h->dVecStaging.alloc(n));
		T *staging = reinterpret_cast<T *>(h->dVecStaging.p);
		convertBackKernel<T><<<1, 128, 0, h->stream>>>(src, staging, n);
		CUDA_TRY(cudaMemcpyAsync(dst, staging, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
		CUDA_TRY(cudaStreamSynchronize(h->stream));
		return PFFRG_OK;
	}

	// ---- exchange over peer memory ------------------------------------------------------------------------------------------
	PushTargets pushTargets(const pffrg_context *h, int buffer, size_t offset, bool includeSelf)
	{
		PushTargets T; T.n = 0;
		if (includeSelf) T.dst[T.n++] = h->peerV4[buffer][h->rank] + offset;
		for (int r = 0; r < h->nRanks; ++r) if (r != h->rank) T.dst[T.n++] = h->peerV4[buffer][r] + offset;
		return T;
	}
	// all ranks have reached this point of their streams (and what they stored into peer memory before it has arrived)
	int peerBarrier(pffrg_context *h, double ms)
	{
		PeerSyncs S; S.n = h->nRanks;
		for (int r = 0; r < h->nRanks; ++r) S.block[r] = h->peerSync[r];
		++h->epoch;
		signalPeersKernel<<<1, 32, 0, h->stream>>>(S, h->rank, h->epoch, ms);
		CUDA_TRY(cudaGetLastError());
		long long timeoutClocks = 60ll * 2000000000ll; // ~60 s at 2 GHz
		if (const char *e = getenv("PFFRG_PEER_TIMEOUT_S")) timeoutClocks = std::max(1ll, atoll(e)) * 2000000000ll;
		waitPeersKernel<<<1, 32, 0, h->stream>>>(h->dSync.p, h->nRanks, h->epoch, timeoutClocks, h->hFlags);
		CUDA_TRY(cudaGetLastError());
		return PFFRG_OK;
	}
	int checkPeerTimeout(pffrg_context *h)
	{
		if (h->hFlags && h->hFlags[0]) return fail(PFFRG_ERR_NCCL, "rank %d: a peer rank did not reach the exchange in time (PFFRG_PEER_TIMEOUT_S)", h->rank);
		return PFFRG_OK;
	}

	// Map the state buffers and sync blocks of all ranks (CUDA IPC; ranks living in this process are addressed directly). Collective.
	// Any failure on any rank leaves ALL ranks on the NCCL exchange.
	int setupPeerExchange(pffrg_context *h)
	{
		struct Blob { long long pid; void *ptr[3]; cudaIpcMemHandle_t handle[3]; int ok; int pad; };
		const int n = h->nRanks;
		if (n > MAX_RANKS) return PFFRG_OK;
		if (const char *e = getenv("PFFRG_EXCHANGE")) if (std::string(e) == "nccl") return PFFRG_OK;
		Blob mine; memset(&mine, 0, sizeof mine);
		mine.pid = (long long)getpid(); mine.ok = 1;
		if (h->dV4b.alloc(h->v4Elements()) != cudaSuccess || h->dSync.alloc(1) != cudaSuccess || cudaMemset(h->dSync.p, 0, sizeof(SyncBlock)) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
		if (mine.ok && !h->hFlags) { if (cudaMallocHost(&h->hFlags, 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; } else h->hFlags[0] = h->hFlags[1] = 0; }
		mine.ptr[0] = h->dV4.p; mine.ptr[1] = h->dV4b.p; mine.ptr[2] = h->dSync.p;
		for (int k = 0; k < 3 && mine.ok; ++k) if (cudaIpcGetMemHandle(&mine.handle[k], mine.ptr[k]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
		DeviceArray<char> dBlobs; std::vector<Blob> all(n);
		CUDA_TRY(dBlobs.alloc(sizeof(Blob) * (n + 1)));
		CUDA_TRY(cudaMemcpyAsync(dBlobs.p + sizeof(Blob) * n, &mine, sizeof(Blob), cudaMemcpyHostToDevice, h->stream));
		NCCL_TRY(nccl().AllGather(dBlobs.p + sizeof(Blob) * n, dBlobs.p, sizeof(Blob), ncclChar, h->comm, h->stream));
		CUDA_TRY(cudaMemcpyAsync(all.data(), dBlobs.p, sizeof(Blob) * n, cudaMemcpyDeviceToHost, h->stream));
		CUDA_TRY(cudaStreamSynchronize(h->stream));
		int ok = 1;
		for (int r = 0; r < n; ++r) ok &= all[r].ok;
		for (int r = 0; r < n && ok; ++r)
		{
			void *p[3] = { nullptr, nullptr, nullptr };
			if (r == h->rank) { for (int k = 0; k < 3; ++k) p[k] = mine.ptr[k]; }
			else if (all[r].pid == mine.pid)
			{
				// another handle of this process: direct peer access
				cudaPointerAttributes attr;
				if (cudaPointerGetAttributes(&attr, all[r].ptr[0]) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
				if (attr.device != h->device) { const cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0; cudaGetLastError(); }
				for (int k = 0; k < 3; ++k) p[k] = all[r].ptr[k];
			}
			else
				for (int k = 0; k < 3 && ok; ++k)
				{
					if (cudaIpcOpenMemHandle(&p[k], all[r].handle[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
					else h->ipcOpened[h->nIpcOpened++] = p[k];
				}
			h->peerV4[0][r] = static_cast<double *>(p[0]); h->peerV4[1][r] = static_cast<double *>(p[1]); h->peerSync[r] = static_cast<SyncBlock *>(p[2]);
		}
		// collective decision
		DeviceArray<int> dOk; CUDA_TRY(dOk.alloc(1));
		CUDA_TRY(cudaMemcpyAsync(dOk.p, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
		NCCL_TRY(nccl().AllReduce(dOk.p, dOk.p, 1, ncclInt, ncclMin, h->comm, h->stream));
		CUDA_TRY(cudaMemcpyAsync(&ok, dOk.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		CUDA_TRY(cudaStreamSynchronize(h->stream));
		dBlobs.release(); dOk.release();
		h->p2p = ok != 0;
		if (!h->p2p)
		{
			for (int k = 0; k < h->nIpcOpened; ++k) cudaIpcCloseMemHandle(h->ipcOpened[k]);
			h->nIpcOpened = 0; cudaGetLastError();
			h->dV4b.release(); h->cur = 0;
			if (getenv("PFFRG_JIT_VERBOSE")) fprintf(stderr, "[pffrg] rank %d: peer-memory exchange unavailable, using the NCCL exchange\n", h->rank);
		}
		return PFFRG_OK;
	}

	float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }
}

namespace
{
	// results of the last finalize_step that were left on the stream: event timings, the ranks' kernel times, peer timeouts
	int collectFinalize(pffrg_context *h)
	{
		if (!h->finalizePending) return PFFRG_OK;
		CUDA_TRY(cudaEventSynchronize(h->ev[6]));
		h->stats.ms_finalize = elapsed(h->ev[4], h->ev[6]);
		h->stats.ms_exchange = elapsed(h->ev[5], h->ev[6]);
		if (h->timesPending) h->rankTimes.assign(h->hTimes, h->hTimes + h->nRanks);
		h->finalizePending = false; h->timesPending = false;
		return checkPeerTimeout(h);
	}
}

extern "C" {

int pffrg_abi_version(void) { return PFFRG_ABI_VERSION; }
const char *pffrg_last_error(void) { return g_lastError.c_str(); }

int pffrg_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

namespace { thread_local bool g_jitSetupFailed = false; }

// allowJit = false: the precompiled kernels only (the second attempt of pffrg_create after a failed run-time compilation)
static int createHandle(const pffrg_desc *d, pffrg_handle *out, bool allowJit)
{
	g_jitSetupFailed = false;
	if (!d || !out) return fail(PFFRG_ERR_ARGUMENT, "null descriptor or output pointer");
	*out = nullptr;
	if (d->abi_version != PFFRG_ABI_VERSION) return fail(PFFRG_ERR_ARGUMENT, "ABI version mismatch: caller %d, library %d", d->abi_version, PFFRG_ABI_VERSION);
	if (d->core < 0 || d->core > 2) return fail(PFFRG_ERR_ARGUMENT, "unknown core %d", d->core);
	if (d->n_frequencies < 2) return fail(PFFRG_ERR_ARGUMENT, "FrequencyDiscretization must contain at least two frequency values");
	if (d->n_sites < 1 || d->n_range < 1) return fail(PFFRG_ERR_ARGUMENT, "empty lattice");
	if (!d->frequencies || !d->sites_rid || !d->sites_perm || !d->inverted_rid || !d->inverted_perm || !d->overlap_offsets || !d->overlap_rid1 || !d->overlap_rid2 || !d->overlap_perm1 || !d->overlap_perm2 || !d->range_fwd_rid || !d->range_inv_rid)
		return fail(PFFRG_ERR_ARGUMENT, "null table pointer in descriptor");
	for (int i = 0; i < d->n_frequencies; ++i)
		if (!(d->frequencies[i] > 0) || (i > 0 && !(d->frequencies[i] > d->frequencies[i - 1]))) return fail(PFFRG_ERR_ARGUMENT, "frequency mesh must be positive and strictly ascending (index %d)", i);
	if (d->n_frequencies > 512) return fail(PFFRG_ERR_UNSUPPORTED, "more than 512 frequencies are not supported");
	if (d->n_sites > 256) return fail(PFFRG_ERR_UNSUPPORTED, "more than 256 representative sites are not supported yet (got %d)", d->n_sites);
	const int L = d->n_sites;
	for (int j = 0; j < L; ++j)
		if (d->sites_rid[j] < 0 || d->sites_rid[j] >= L || d->inverted_rid[j] < 0 || d->inverted_rid[j] >= L) return fail(PFFRG_ERR_ARGUMENT, "site table entry %d out of range", j);
	if (d->overlap_offsets[0] != 0) return fail(PFFRG_ERR_ARGUMENT, "overlap_offsets[0] must be 0");
	for (int r = 0; r < L; ++r) if (d->overlap_offsets[r + 1] < d->overlap_offsets[r]) return fail(PFFRG_ERR_ARGUMENT, "overlap_offsets not monotone");
	for (int i = 0; i < d->overlap_offsets[L]; ++i)
		if (d->overlap_rid1[i] < 0 || d->overlap_rid1[i] >= L || d->overlap_rid2[i] < 0 || d->overlap_rid2[i] >= L) return fail(PFFRG_ERR_ARGUMENT, "overlap entry %d out of range", i);
	for (int j = 0; j < d->n_range; ++j)
		if (d->range_fwd_rid[j] < 0 || d->range_fwd_rid[j] >= L || d->range_inv_rid[j] < 0 || d->range_inv_rid[j] >= L) return fail(PFFRG_ERR_ARGUMENT, "range entry %d out of range", j);

	int nDev = pffrg_device_count();
	if (nDev <= 0) return fail(PFFRG_ERR_CUDA, "no CUDA device available (libpffrg has no CPU fallback)");
	if (d->device < 0 || d->device >= nDev) return fail(PFFRG_ERR_ARGUMENT, "device %d out of range (%d devices)", d->device, nDev);
	CUDA_TRY(cudaSetDevice(d->device));
	cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
	if (prop.major < 10) return fail(PFFRG_ERR_CUDA, "device %d is sm_%d%d; libpffrg is built for sm_100a only", d->device, prop.major, prop.minor);
	const int smCount = prop.multiProcessorCount;

	const pffrg_desc *original = d;
	const RelabelledDesc relabelled(original);
	d = &relabelled.view; // from here on: the device-internal site order
	pffrg_context *h = new pffrg_context();
	if (!relabelled.identity) { h->siteNewOf = relabelled.newOf; h->siteOrder = relabelled.order; }
	const CoreModel m = modelOf(d->core);
	h->core = d->core; h->nw = d->n_frequencies; h->L = L; h->Lp = paddedSites(L); h->C = m.C; h->RL = m.C * h->Lp; h->nArrays = m.arrays;
	h->nf = (int64_t)h->nw * h->nw * (h->nw + 1) / 2;
	h->device = d->device; h->spin = d->spin_length; h->smCount = smCount;
	if (const char *e = getenv("PFFRG_PERSISTENT")) h->persistent = atoi(e) != 0;
	h->mesh.assign(d->frequencies, d->frequencies + h->nw);
	h->overlapTotal = d->overlap_offsets[L];
	if ((double)h->nf * h->RL > 2.0e9) { delete h; return fail(PFFRG_ERR_UNSUPPORTED, "vertex too large for 32-bit row offsets"); }

	// launch configuration: k groups of L threads; batch width NB chosen so that at least two CTAs fit per SM
	// a group of threads covers the L sites of one quadrature node; groups are padded to whole warps when that idles at
	// most a quarter of the lanes (then every warp gathers from one node only: fewer cache lines per load, uniform table reads)
	const char *jitEnv = getenv("PFFRG_JIT");
	const bool jitEnabled = allowJit && !(jitEnv && atoi(jitEnv) == 0);
	const LaunchGeometry geo = chooseGeometry(L, d->core == SU2 && jitEnabled);
	h->stride = geo.stride; h->groups = geo.groups; h->threads = geo.threads;
	// SU2/XYZ: two CTAs per SM (100 KB each); the TRI core stages four 16-channel RPA operand buffers and runs one CTA per SM
	h->nb = 32;
	const int minNb = h->core == TRI ? 4 : 8;
	const size_t smemTarget = h->core == TRI ? 200 * 1024 : 100 * 1024;
	while (h->nb > minNb && flowSmemBytes(h->core, h->nb, h->nw, L, h->groups) > smemTarget) h->nb >>= 1;
	// TRI core with the Gram form of the RPA phase (run-time compiled, rpaTriGram): gather batch = staged nodes = 8
	if (h->core == TRI && wantTriGram(h->core) && jitEnabled && 16 * L <= 1024) h->nb = 8;
	// PFFRG_NB: force the gather batch of the precompiled kernels (tests exercise every kernel variant on small lattices)
	if (const char *e = getenv("PFFRG_NB")) { const int v = atoi(e); if ((v == 32 || v == 16 || v == 8 || (v == 4 && h->core == TRI)) && v <= h->nb) h->nb = v; }
	h->smemBytes = flowSmemBytes(h->core, h->nb, h->nw, L, h->groups);
	if (h->smemBytes > (size_t)prop.sharedMemPerBlockOptin) { const size_t need = h->smemBytes; delete h; return fail(PFFRG_ERR_UNSUPPORTED, "flow kernel needs %zu bytes of shared memory", need); }
	const bool tri8 = h->core == TRI && h->nb == 8; // rpaTri8: one task stream per warp, staging layout with TRI8_RID_STRIDE
	h->nslots = tri8 ? h->threads / 32 : (h->threads / 32) * (32 / h->nb);

	std::vector<unsigned> words; std::vector<int4> tasks; std::vector<int> slotOff;
	if (tri8) buildRpaTri8(d, h->nslots, words, tasks, slotOff, h->uniquePairs);
	else buildRpa(d, h->core, h->nslots, h->nb + 1, words, tasks, slotOff, h->uniquePairs);

	std::vector<int> sitesPerm(L), invPerm(L);
	for (int j = 0; j < L; ++j) { sitesPerm[j] = packPerm(d->sites_perm + 3 * j); invPerm[j] = packPerm(d->inverted_perm + 3 * j); }

	cudaError_t e = cudaSuccess;
	auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
	ok(h->dMesh.upload(h->mesh));
	ok(h->dMeshStart.upload(buildMeshIndex(h->mesh, h->meshShift, h->meshKeyBase, h->meshKeys)));
	ok(h->dSitesRid.upload(std::vector<int>(d->sites_rid, d->sites_rid + L)));
	ok(h->dInvRid.upload(std::vector<int>(d->inverted_rid, d->inverted_rid + L)));
	ok(h->dSitesPerm.upload(sitesPerm)); ok(h->dInvPerm.upload(invPerm));
	ok(h->dRngFwd.upload(std::vector<int>(d->range_fwd_rid, d->range_fwd_rid + d->n_range)));
	ok(h->dRngInv.upload(std::vector<int>(d->range_inv_rid, d->range_inv_rid + d->n_range)));
	ok(h->dTasks.upload(tasks)); ok(h->dSlotOff.upload(slotOff)); ok(h->dWords.upload(words));
	if (!h->siteNewOf.empty()) ok(h->dSiteNewOf.upload(h->siteNewOf));
	ok(h->dV4.alloc(h->v4Elements())); ok(h->dFlow4.alloc(h->v4Elements()));
	ok(h->dV2.alloc(h->nw)); ok(h->dFlow2.alloc(h->nw)); ok(h->dCutoff.alloc(1));
	h->nodeStride = 2 * h->nw + 8;
	ok(h->dCount.alloc(h->nw)); ok(h->dNodeW.alloc((size_t)h->nw * h->nodeStride)); ok(h->dNodeWt.alloc((size_t)h->nw * h->nodeStride));
	ok(h->dNan.alloc(1));
	if (e == cudaSuccess) e = cudaMemset(h->dV4.p, 0, h->v4Elements() * sizeof(double));
	if (e == cudaSuccess) e = cudaMemset(h->dFlow4.p, 0, h->v4Elements() * sizeof(double));
	if (e == cudaSuccess) e = cudaMemset(h->dFlow2.p, 0, h->nw * sizeof(double));
	if (e == cudaSuccess) e = cudaMemset(h->dNan.p, 0, sizeof(int));
	if (e == cudaSuccess) e = cudaMallocHost(&h->hNan, sizeof(int));
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
	for (auto &ev : h->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
	if (e != cudaSuccess)
	{
		int code = fail(PFFRG_ERR_CUDA, "device setup failed: %s", cudaGetErrorString(e));
		pffrg_destroy(h);
		return code;
	}
	h->bounds = { 0, h->nf };
	if (const char *e = getenv("PFFRG_ORDER")) h->itemOrder = (e[0] == 't' || e[0] == '1') ? 1 : 0;
	const int jitStatus = setupJit(h, d, (size_t)prop.sharedMemPerBlockOptin, jitEnabled);
	if (jitStatus != PFFRG_OK) { pffrg_destroy(h); g_jitSetupFailed = true; return jitStatus; }
	*out = h;
	return PFFRG_OK;
}

// Run-time compilation can fail for reasons that have nothing to do with the lattice (no NVRTC at run time, a compiler error on another toolkit,
// a full disk): the precompiled kernels can still run it, slower. So a failed setup of the run-time compiled kernel is reported on stderr and the
// handle is built again without it -- unless the caller asked for a specific kernel form (PFFRG_RPA, PFFRG_SPLIT, ...) or for PFFRG_JIT_STRICT=1,
// where the error stays an error.
int pffrg_create(const pffrg_desc *d, pffrg_handle *out)
{
	const int rc = createHandle(d, out, true);
	if (rc == PFFRG_OK || !g_jitSetupFailed) return rc;
	const char *strict = getenv("PFFRG_JIT_STRICT");
	if ((strict && atoi(strict) != 0) || getenv("PFFRG_RPA") || getenv("PFFRG_SPLIT") || getenv("PFFRG_SUBCTAS") || getenv("PFFRG_JIT_NBT")) return rc;
	fprintf(stderr, "[pffrg] %s -- continuing with the precompiled kernels (PFFRG_JIT_STRICT=1 makes this an error)\n", pffrg_last_error());
	return createHandle(d, out, false);
}

int pffrg_destroy(pffrg_handle h)
{
	if (!h) return PFFRG_OK;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	for (int k = 0; k < h->nIpcOpened; ++k) cudaIpcCloseMemHandle(h->ipcOpened[k]);
	if (h->comm) nccl().CommDestroy(h->comm);
	h->dV4b.release(); h->dSync.release(); h->dVecStaging.release(); h->dTimes.release();
	if (h->hFlags) cudaFreeHost(h->hFlags);
	h->dMesh.release(); h->dSitesRid.release(); h->dInvRid.release(); h->dSitesPerm.release(); h->dInvPerm.release(); h->dRngFwd.release(); h->dRngInv.release();
	h->dSlotOff.release(); h->dTasks.release(); h->dWords.release(); h->dMeshStart.release(); h->dGramTerms.release(); h->dGramSeg.release(); h->dSiteNewOf.release(); h->dTriBlocks.release();
	if (h->jitLibrary) cudaLibraryUnload(h->jitLibrary); h->dV4.release(); h->dFlow4.release(); h->dV2.release(); h->dFlow2.release(); h->dCutoff.release();
	h->dCount.release(); h->dNodeW.release(); h->dNodeWt.release(); h->dNan.release(); h->dStaging.release();
	h->dChiPartial.release(); h->dChi.release(); h->dChiCount.release();
	if (h->hNan) cudaFreeHost(h->hNan);
	if (h->hTimes) cudaFreeHost(h->hTimes);
	for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
	return PFFRG_OK;
}

int pffrg_num_vertex_arrays(pffrg_handle h) { return h ? h->nArrays : fail(PFFRG_ERR_ARGUMENT, "null handle"); }
int64_t pffrg_vertex_array_length(pffrg_handle h) { return h ? h->nf * h->L * (h->core == TRI ? 16 : 1) : fail(PFFRG_ERR_ARGUMENT, "null handle"); }
int64_t pffrg_num_items(pffrg_handle h) { return h ? h->nf : fail(PFFRG_ERR_ARGUMENT, "null handle"); }

int pffrg_comm_unique_id(void *idOut)
{
	if (!idOut) return fail(PFFRG_ERR_ARGUMENT, "null id buffer");
	static_assert(sizeof(ncclUniqueId) <= PFFRG_UNIQUE_ID_BYTES, "unique id does not fit");
	if (!nccl().ok) return fail(PFFRG_ERR_NCCL, "%s", nccl().error.c_str());
	ncclUniqueId id;
	NCCL_TRY(nccl().GetUniqueId(&id));
	memset(idOut, 0, PFFRG_UNIQUE_ID_BYTES);
	memcpy(idOut, &id, sizeof(id));
	return PFFRG_OK;
}

int pffrg_comm_init(pffrg_handle h, const void *id, int rank, int nRanks)
{
	if (!h || !id) return fail(PFFRG_ERR_ARGUMENT, "null handle or id");
	if (nRanks < 1 || rank < 0 || rank >= nRanks) return fail(PFFRG_ERR_ARGUMENT, "bad rank %d of %d", rank, nRanks);
	if (h->comm) return fail(PFFRG_ERR_STATE, "communicator already initialised");
	CUDA_TRY(cudaSetDevice(h->device));
	ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
	if (!nccl().ok) return fail(PFFRG_ERR_NCCL, "%s", nccl().error.c_str());
	NCCL_TRY(nccl().CommInitRank(&h->comm, nRanks, uid, rank));
	if (const char *e = getenv("PFFRG_BALANCE")) h->balance = atoi(e) != 0;
	h->rankTimes.clear();
	h->rank = rank; h->nRanks = nRanks;
	h->bounds.assign(nRanks + 1, 0); h->bounds[nRanks] = h->nf;
	if (nRanks > 1) { const int rc = setupPeerExchange(h); if (rc != PFFRG_OK) return rc; }
	return PFFRG_OK;
}

int pffrg_item_range(pffrg_handle h, int64_t *begin, int64_t *end)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (begin) *begin = h->curBegin;
	if (end) *end = h->curEnd;
	return PFFRG_OK;
}

int pffrg_set_item_range(pffrg_handle h, int64_t begin, int64_t end)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (end > begin && (begin < 0 || end > h->nf)) return fail(PFFRG_ERR_ARGUMENT, "item range [%lld, %lld) outside [0, %lld)", (long long)begin, (long long)end, (long long)h->nf);
	h->userBegin = begin; h->userEnd = end;
	return PFFRG_OK;
}

int pffrg_set_state(pffrg_handle h, double cutoff, const void *v2, const void *const *v4, int dtype)
{
	if (!h || !v2 || !v4) return fail(PFFRG_ERR_ARGUMENT, "null argument");
	if (dtype != PFFRG_F32 && dtype != PFFRG_F64) return fail(PFFRG_ERR_ARGUMENT, "unknown dtype %d", dtype);
	CUDA_TRY(cudaSetDevice(h->device));
	int rc = dtype == PFFRG_F64 ? importArrays<double>(h, v4, h->v4cur(), 0, h->nf) : importArrays<float>(h, v4, h->v4cur(), 0, h->nf);
	if (rc != PFFRG_OK) return rc;
	rc = dtype == PFFRG_F64 ? importVector<double>(h, v2, h->dV2.p, h->nw) : importVector<float>(h, v2, h->dV2.p, h->nw);
	if (rc != PFFRG_OK) return rc;
	h->cutoff = cutoff;
	setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, cutoff);
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	h->haveState = true; h->haveFlow = false;
	return PFFRG_OK;
}

int pffrg_upload_slice(pffrg_handle h, int64_t *begin, int64_t *end)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (begin) *begin = h->nf * h->rank / h->nRanks;
	if (end) *end = h->nf * (h->rank + 1) / h->nRanks;
	return PFFRG_OK;
}

int pffrg_set_state_sharded(pffrg_handle h, double cutoff, const void *v2, const void *const *v4, int dtype)
{
	if (!h || !v2 || !v4) return fail(PFFRG_ERR_ARGUMENT, "null argument");
	if (h->nRanks <= 1 || !h->p2p) return pffrg_set_state(h, cutoff, v2, v4, dtype);
	if (dtype != PFFRG_F32 && dtype != PFFRG_F64) return fail(PFFRG_ERR_ARGUMENT, "unknown dtype %d", dtype);
	CUDA_TRY(cudaSetDevice(h->device));
	{ const int rc = collectFinalize(h); if (rc != PFFRG_OK) return rc; }
	int64_t begin = 0, end = 0; pffrg_upload_slice(h, &begin, &end);
	// every rank is done reading the buffer that is about to be overwritten
	int rc = peerBarrier(h, 0.0);
	if (rc != PFFRG_OK) return rc;
	rc = dtype == PFFRG_F64 ? importArrays<double>(h, v4, h->v4cur(), begin, end - begin) : importArrays<float>(h, v4, h->v4cur(), begin, end - begin);
	if (rc != PFFRG_OK) return rc;
	const size_t off = (size_t)begin * h->RL, cnt = (size_t)(end - begin) * h->RL;
	if (cnt > 0)
	{
		const unsigned blocks = (unsigned)std::min<size_t>(1184, (cnt / 2 + 255) / 256);
		copyPushKernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(h->v4cur() + off), cnt / 2, pushTargets(h, h->cur, off, false));
		CUDA_TRY(cudaGetLastError());
	}
	rc = peerBarrier(h, 0.0);
	if (rc != PFFRG_OK) return rc;
	rc = dtype == PFFRG_F64 ? importVector<double>(h, v2, h->dV2.p, h->nw) : importVector<float>(h, v2, h->dV2.p, h->nw);
	if (rc != PFFRG_OK) return rc;
	h->cutoff = cutoff;
	setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, cutoff);
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	h->haveState = true; h->haveFlow = false;
	return checkPeerTimeout(h);
}

int pffrg_get_state_slice(pffrg_handle h, double *cutoff, void *v2, void *const *v4, int dtype, int64_t begin, int64_t end)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (!h->haveState) return fail(PFFRG_ERR_STATE, "no state has been set");
	if (dtype != PFFRG_F32 && dtype != PFFRG_F64) return fail(PFFRG_ERR_ARGUMENT, "unknown dtype %d", dtype);
	if (begin < 0 || end > h->nf || end < begin) return fail(PFFRG_ERR_ARGUMENT, "item range [%lld, %lld) outside [0, %lld)", (long long)begin, (long long)end, (long long)h->nf);
	CUDA_TRY(cudaSetDevice(h->device));
	if (cutoff) *cutoff = h->cutoff;
	if (v2) { int rc = dtype == PFFRG_F64 ? exportVector<double>(h, h->dV2.p, v2, h->nw) : exportVector<float>(h, h->dV2.p, v2, h->nw); if (rc != PFFRG_OK) return rc; }
	if (v4) { int rc = dtype == PFFRG_F64 ? exportArrays<double>(h, h->v4cur(), v4, begin, end - begin) : exportArrays<float>(h, h->v4cur(), v4, begin, end - begin); if (rc != PFFRG_OK) return rc; }
	return PFFRG_OK;
}

int pffrg_set_initial_condition(pffrg_handle h, double cutoff, const double *bare)
{
	if (!h || !bare) return fail(PFFRG_ERR_ARGUMENT, "null argument");
	CUDA_TRY(cudaSetDevice(h->device));
	const size_t entries = (size_t)h->C * h->L;
	if (h->dStaging.n < entries) CUDA_TRY(h->dStaging.alloc(entries));
	std::vector<double> ordered;
	if (!h->siteOrder.empty())
	{
		ordered.resize(entries);
		for (int c = 0; c < h->C; ++c) for (int n = 0; n < h->L; ++n) ordered[(size_t)c * h->L + n] = bare[(size_t)c * h->L + h->siteOrder[n]];
		bare = ordered.data();
	}
	CUDA_TRY(cudaMemcpyAsync(h->dStaging.p, bare, entries * sizeof(double), cudaMemcpyHostToDevice, h->stream));
	initialConditionKernel<<<1184, 256, 0, h->stream>>>(h->v4cur(), h->dStaging.p, (size_t)h->nf, h->L, h->Lp, h->RL, vectorWidth(h->core));
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemsetAsync(h->dV2.p, 0, h->nw * sizeof(double), h->stream));
	h->cutoff = cutoff;
	setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, cutoff);
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	h->haveState = true; h->haveFlow = false;
	return PFFRG_OK;
}

int pffrg_get_state(pffrg_handle h, double *cutoff, void *v2, void *const *v4, int dtype)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (!h->haveState) return fail(PFFRG_ERR_STATE, "no state has been set");
	if (dtype != PFFRG_F32 && dtype != PFFRG_F64) return fail(PFFRG_ERR_ARGUMENT, "unknown dtype %d", dtype);
	CUDA_TRY(cudaSetDevice(h->device));
	if (cutoff) *cutoff = h->cutoff;
	if (v2) { int rc = dtype == PFFRG_F64 ? exportVector<double>(h, h->dV2.p, v2, h->nw) : exportVector<float>(h, h->dV2.p, v2, h->nw); if (rc != PFFRG_OK) return rc; }
	if (v4) { int rc = dtype == PFFRG_F64 ? exportArrays<double>(h, h->v4cur(), v4, 0, h->nf) : exportArrays<float>(h, h->v4cur(), v4, 0, h->nf); if (rc != PFFRG_OK) return rc; }
	return PFFRG_OK;
}

int pffrg_get_flow(pffrg_handle h, void *v2flow, void *const *v4flow, int dtype)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (!h->haveFlow) return fail(PFFRG_ERR_STATE, "no flow has been computed");
	if (dtype != PFFRG_F32 && dtype != PFFRG_F64) return fail(PFFRG_ERR_ARGUMENT, "unknown dtype %d", dtype);
	CUDA_TRY(cudaSetDevice(h->device));
	if (!h->flowGathered)
	{
		int rc = exchangeSlices(h, h->dFlow4.p);
		if (rc != PFFRG_OK) return rc;
		h->flowGathered = true;
	}
	if (v2flow) { int rc = dtype == PFFRG_F64 ? exportVector<double>(h, h->dFlow2.p, v2flow, h->nw) : exportVector<float>(h, h->dFlow2.p, v2flow, h->nw); if (rc != PFFRG_OK) return rc; }
	if (v4flow) { int rc = dtype == PFFRG_F64 ? exportArrays<double>(h, h->dFlow4.p, v4flow, 0, h->nf) : exportArrays<float>(h, h->dFlow4.p, v4flow, 0, h->nf); if (rc != PFFRG_OK) return rc; }
	return PFFRG_OK;
}

int pffrg_compute_step(pffrg_handle h, int *diverged)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (!h->haveState) return fail(PFFRG_ERR_STATE, "compute_step before set_state");
	CUDA_TRY(cudaSetDevice(h->device));
	{ const int rc = collectFinalize(h); if (rc != PFFRG_OK) return rc; }
	const std::vector<int> counts = hostNodeCounts(h);
	partitionItems(h, counts);
	int64_t begin = h->bounds[h->rank], end = h->bounds[h->rank + 1];
	if (h->userEnd > h->userBegin) { begin = h->userBegin; end = h->userEnd; }
	h->curBegin = begin; h->curEnd = end;

	const Problem P = h->problem();
	CUDA_TRY(cudaMemsetAsync(h->dNan.p, 0, sizeof(int), h->stream));
	CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
	const size_t smemV2 = sizeof(double) * (h->nw + 128);
	if (h->core == SU2) v2FlowKernel<SU2><<<h->nw, 128, smemV2, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->dFlow2.p);
	else if (h->core == XYZ) v2FlowKernel<XYZ><<<h->nw, 128, smemV2, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->dFlow2.p);
	else v2FlowKernel<TRI><<<h->nw, 128, smemV2, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->dFlow2.p);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
	nodeTableKernel<<<h->nw, 128, sizeof(double) * (3 * h->nw + 2 * h->nodeStride), h->stream>>>(P, h->nodeTable(), h->dV2.p, h->dFlow2.p, h->dCutoff.p);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
	CUDA_TRY(launchFlowDispatch(h, begin, end - begin));
	CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
	h->stats.launches = 3;
	if (h->nRanks > 1 && !(h->userEnd > h->userBegin)) NCCL_TRY(nccl().AllReduce(h->dNan.p, h->dNan.p, 1, ncclInt, ncclMax, h->comm, h->stream));
	CUDA_TRY(cudaMemcpyAsync(h->hNan, h->dNan.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	fillStats(h, counts, begin, end); // host work overlaps with the kernels
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	h->stats.ms_v2_flow = elapsed(h->ev[0], h->ev[1]);
	h->stats.ms_node_table = elapsed(h->ev[1], h->ev[2]);
	h->stats.ms_v4_flow = elapsed(h->ev[2], h->ev[3]);
	h->haveFlow = true;
	h->flowGathered = (h->nRanks <= 1);
	if (diverged) *diverged = *h->hNan ? 1 : 0;
	return checkPeerTimeout(h);
}

int pffrg_finalize_step(pffrg_handle h, double newCutoff)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	if (!h->haveFlow) return fail(PFFRG_ERR_STATE, "finalize_step before compute_step");
	CUDA_TRY(cudaSetDevice(h->device));
	CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
	// every rank updates the self energy (its flow is computed redundantly) and its own slice of the vertex
	eulerKernel<<<1, 128, 0, h->stream>>>(h->dV2.p, h->dFlow2.p, (size_t)h->nw, h->dCutoff.p, newCutoff);
	CUDA_TRY(cudaGetLastError());
	const size_t off = (size_t)h->curBegin * h->RL, cnt = (size_t)(h->curEnd - h->curBegin) * h->RL;
	const bool sharded = h->nRanks > 1 && !(h->userEnd > h->userBegin);
	h->stats.launches += 2;
	if (sharded && h->p2p)
	{
		// Euler update of the own slice + its distribution in one kernel: the new values go into the slice of EVERY rank's next-state
		// buffer over NVLink (the ranks still read the old state until they have all arrived)
		if (cnt > 0)
		{
			const unsigned blocks = (unsigned)std::min<size_t>(1184, (cnt / 2 + 255) / 256);
			eulerPushKernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(h->v4cur() + off), reinterpret_cast<const double2 *>(h->dFlow4.p + off), cnt / 2, h->dCutoff.p, newCutoff, pushTargets(h, 1 - h->cur, off, true));
			CUDA_TRY(cudaGetLastError());
			++h->stats.launches;
		}
		setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, newCutoff);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(h->ev[5], h->stream));
		{ const int rc = peerBarrier(h, (double)h->stats.ms_v4_flow); if (rc != PFFRG_OK) return rc; }
		h->stats.launches += 2;
		h->cur ^= 1;
		if (h->balance)
		{
			if (!h->hTimes) CUDA_TRY(cudaMallocHost(&h->hTimes, sizeof(double) * MAX_RANKS));
			CUDA_TRY(cudaMemcpyAsync(h->hTimes, h->dSync.p->times, sizeof(double) * h->nRanks, cudaMemcpyDeviceToHost, h->stream));
		}
	}
	else
	{
		if (cnt > 0)
		{
			eulerKernel<<<1184, 256, 0, h->stream>>>(h->v4cur() + off, h->dFlow4.p + off, cnt, h->dCutoff.p, newCutoff);
			CUDA_TRY(cudaGetLastError());
			++h->stats.launches;
		}
		setScalarKernel<<<1, 1, 0, h->stream>>>(h->dCutoff.p, newCutoff);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaEventRecord(h->ev[5], h->stream));
		if (sharded) { int rc = exchangeSlices(h, h->v4cur()); if (rc != PFFRG_OK) return rc; }
		if (sharded && h->balance)
		{
			// every rank's flow-kernel time of this step -> all ranks (a sum over vectors with one non-zero entry each)
			if (!h->hTimes) { CUDA_TRY(cudaMallocHost(&h->hTimes, sizeof(double) * MAX_RANKS)); }
			if (!h->dTimes.p) CUDA_TRY(h->dTimes.alloc(MAX_RANKS));
			for (int r = 0; r < h->nRanks; ++r) h->hTimes[r] = r == h->rank ? (double)h->stats.ms_v4_flow : 0.0;
			CUDA_TRY(cudaMemcpyAsync(h->dTimes.p, h->hTimes, sizeof(double) * h->nRanks, cudaMemcpyHostToDevice, h->stream));
			NCCL_TRY(nccl().AllReduce(h->dTimes.p, h->dTimes.p, h->nRanks, ncclDouble, ncclSum, h->comm, h->stream));
			CUDA_TRY(cudaMemcpyAsync(h->hTimes, h->dTimes.p, sizeof(double) * h->nRanks, cudaMemcpyDeviceToHost, h->stream));
		}
	}
	CUDA_TRY(cudaEventRecord(h->ev[6], h->stream));
	// no host synchronisation here: the times (and the event timings) are collected by the next call that needs them
	h->finalizePending = true;
	h->timesPending = sharded && h->balance;
	h->cutoff = newCutoff;
	h->haveFlow = false;
	return PFFRG_OK;
}

int pffrg_num_channels(pffrg_handle h) { return h ? h->C : fail(PFFRG_ERR_ARGUMENT, "null handle"); }

int pffrg_measure_correlation(pffrg_handle h, double *chi)
{
	if (!h || !chi) return fail(PFFRG_ERR_ARGUMENT, "null argument");
	if (!h->haveState) return fail(PFFRG_ERR_STATE, "measure_correlation before set_state");
	CUDA_TRY(cudaSetDevice(h->device));
	const int entries = h->C * h->L;
	if (!h->dChi.p)
	{
		CUDA_TRY(h->dChiPartial.alloc((size_t)h->nodeStride * entries));
		CUDA_TRY(h->dChi.alloc(entries));
		CUDA_TRY(h->dChiCount.alloc(1));
	}
	const Problem P = h->problem();
	const int groups = std::max(1, 256 / h->L);
	const size_t smem = sizeof(double) * ((size_t)2 * h->nw + 2 * h->nodeStride + (size_t)groups * entries);
	if (h->core == SU2) { CUDA_TRY(cudaFuncSetAttribute(correlationKernel<SU2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); correlationKernel<SU2><<<h->nodeStride, 256, smem, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->nodeStride, h->dChiPartial.p, h->dChiCount.p); }
	else if (h->core == XYZ) { CUDA_TRY(cudaFuncSetAttribute(correlationKernel<XYZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); correlationKernel<XYZ><<<h->nodeStride, 256, smem, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->nodeStride, h->dChiPartial.p, h->dChiCount.p); }
	else { CUDA_TRY(cudaFuncSetAttribute(correlationKernel<TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); correlationKernel<TRI><<<h->nodeStride, 256, smem, h->stream>>>(P, h->v4cur(), h->dV2.p, h->dCutoff.p, h->nodeStride, h->dChiPartial.p, h->dChiCount.p); }
	CUDA_TRY(cudaGetLastError());
	correlationSumKernel<<<(entries + 127) / 128, 128, 0, h->stream>>>(h->dChiPartial.p, h->dChiCount.p, entries, h->dChi.p);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpyAsync(chi, h->dChi.p, sizeof(double) * entries, cudaMemcpyDeviceToHost, h->stream));
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	if (!h->siteNewOf.empty())
	{
		// device-internal site order -> the reference's
		const std::vector<double> device(chi, chi + entries);
		for (int c = 0; c < h->C; ++c) for (int j = 0; j < h->L; ++j) chi[(size_t)c * h->L + j] = device[(size_t)c * h->L + h->siteNewOf[j]];
	}
	return PFFRG_OK;
}

int pffrg_synchronize(pffrg_handle h)
{
	if (!h) return fail(PFFRG_ERR_ARGUMENT, "null handle");
	CUDA_TRY(cudaSetDevice(h->device));
	CUDA_TRY(cudaStreamSynchronize(h->stream));
	return PFFRG_OK;
}

int pffrg_get_stats(pffrg_handle h, pffrg_stats *out)
{
	if (!h || !out) return fail(PFFRG_ERR_ARGUMENT, "null argument");
	if (h->finalizePending) { CUDA_TRY(cudaSetDevice(h->device)); const int rc = collectFinalize(h); if (rc != PFFRG_OK) return rc; }
	*out = h->stats;
	out->jit_rpa = h->jitKernel ? 1 : 0;
	out->jit_compile_ms = h->jitCompileMs;
	out->threads = h->threads; out->smem_bytes = (int32_t)h->smemBytes; out->node_batch = h->nb; out->rpa_batch = h->jitKernel ? h->nbt * h->subs : h->nb; out->sub_ctas = h->jitKernel ? h->subs : 1;
	out->autotuned_shapes = h->autotuned;
	out->gram_rows = h->gramRows; out->rpa_terms_merged = (int32_t)h->uniquePairs;
	out->gather_threads = h->jitKernel ? h->splitGather : 0; out->producer_warps = h->jitKernel ? h->producerWarps : 0;
	out->rpa_warps = h->jitKernel ? h->rpaWarps : h->threads / 32; out->min_blocks = h->jitKernel ? h->minBlocks : (h->core == TRI ? 1 : 2);
	return PFFRG_OK;
}

void *pffrg_stream(pffrg_handle h) { return h ? (void *)h->stream : nullptr; }

int pffrg_jit_compile_check(const pffrg_desc *d, int64_t *cubinBytes)
{
	if (!d || d->n_sites < 1 || d->n_sites > 256 || d->core < 0 || d->core > 2 || !d->overlap_offsets) return fail(PFFRG_ERR_ARGUMENT, "bad descriptor");
	// same launch configuration as pffrg_create
	const int L = d->n_sites;
	const char *jitEnv = getenv("PFFRG_JIT");
	const LaunchGeometry geo = chooseGeometry(L, d->core == SU2 && !(jitEnv && atoi(jitEnv) == 0));
	const int groups = geo.groups, threads = geo.threads;
	const RelabelledDesc relabelled(d);
	d = &relabelled.view; // as pffrg_create
	int64_t uniquePairs = 0;
	for (int rid = 0; rid < L; ++rid) uniquePairs += (int64_t)mergedOverlap(d, d->core, rid).size();
	if (d->core == TRI)
	{
		if (!wantTriGram(d->core)) return fail(PFFRG_ERR_UNSUPPORTED, "the TRI core has no run-time compiled kernel with PFFRG_RPA=table");
		const int Lp = paddedSites(L);
		const JitShape g = chooseTriGramShape(d->n_frequencies, L, groups, threads, 227 * 1024);
		if (!g.nb) return fail(PFFRG_ERR_UNSUPPORTED, "the TRI Gram form does not fit this lattice");
		std::vector<char> cubin;
		const std::string err = compileFlowKernel(d->core, g.nb, g.nbt, 1, 1, threads, g.minBlocks, KernelSizes{ L, Lp, channelsOf(d->core) * Lp, d->n_frequencies }, std::string(), cubin, triGramDefines(g));
		if (!err.empty()) return fail(PFFRG_ERR_CUDA, "%s", err.c_str());
		std::vector<unsigned short> blocks; std::vector<unsigned> terms; std::vector<int> seg; double conflicts = 0.0;
		buildTriGramTables(d, L, g.gramRows, threads / 32, blocks, terms, seg, &conflicts);
		if (getenv("PFFRG_JIT_VERBOSE")) fprintf(stderr, "[pffrg tri gram] threads %d smem %zu resident blocks %d rounds %zu words %zu bank-conflict degree %.3f\n", threads, g.smem, g.gramRows, blocks.size() / g.gramRows, terms.size(), conflicts);
		if (cubinBytes) *cubinBytes = (int64_t)cubin.size();
		return PFFRG_OK;
	}
	if (wantGram(d->core, uniquePairs))
	{
		const int Lp = paddedSites(L);
		const GramLaunch launch = chooseGramLaunch(d->n_frequencies, L, Lp, groups, geo.stride, threads, 227 * 1024, uniquePairs, !getenv("PFFRG_RPA"), d->core);
		const JitShape &g = launch.shape;
		if (!g.nb && getenv("PFFRG_RPA")) return fail(PFFRG_ERR_UNSUPPORTED, "no launch shape for the Gram form of the RPA phase");
		if (g.nb)
		{
		std::vector<char> cubin;
		const std::string err = compileFlowKernel(d->core, g.nb, g.nbt, 1, 1, launch.threads, g.minBlocks, KernelSizes{ L, Lp, channelsOf(d->core) * Lp, d->n_frequencies }, std::string(), cubin, gramDefines(g, d->core, L, Lp));
		if (!err.empty()) return fail(PFFRG_ERR_CUDA, "%s", err.c_str());
		std::vector<unsigned> terms; std::vector<int> seg; double conflicts = 0.0;
		buildGramTables(d, L, gramGeometry(d->core, L, Lp).lp, g.gramRows, launch.reduceWarps, terms, seg, &conflicts);
		if (getenv("PFFRG_JIT_VERBOSE")) fprintf(stderr, "[pffrg gram] threads %d nb %d nbt %d ctas %d smem %zu rows/block %d gemm threads %d words %zu (merged terms %lld) bank-conflict degree %.3f\n", threads, g.nb, g.nbt, g.minBlocks, g.smem, g.gramRows, g.gramThreads, terms.size(), (long long)uniquePairs, conflicts);
		if (cubinBytes) *cubinBytes = (int64_t)cubin.size();
		return PFFRG_OK;
		}
	}
	JitShape shape = chooseJitShape(d->core, d->n_frequencies, L, groups, threads / 32, 227 * 1024);
	if (!shape.nb) return fail(PFFRG_ERR_UNSUPPORTED, "lattice too large for the specialised kernel");
	int subs = 1;
	if (const char *e = getenv("PFFRG_SUBCTAS")) subs = std::max(1, atoi(e));
	if (subs > 1)
	{
		shape = subCtaShape(d->core, d->n_frequencies, L, groups, threads / 32, subs, shape.nbt, shape.nb, getenv("PFFRG_JIT_MINBLOCKS") ? std::max(1, atoi(getenv("PFFRG_JIT_MINBLOCKS"))) : 1, 227 * 1024);
		if (!shape.nb) return fail(PFFRG_ERR_UNSUPPORTED, "PFFRG_SUBCTAS=%d does not fit (shared memory / warps)", subs);
	}
	shape.cluster = shape.subs > 1 ? 1 : 2; // as in setupJit
	if (const char *e = getenv("PFFRG_CLUSTER")) shape.cluster = std::min(8, std::max(1, atoi(e)));
	RpaProgram prog = buildRpaProgram(d, d->core, shape.nbt * shape.subs, shape.rpaWarps);
	prog.maxAccumulators = defaultAccumulators(threads * shape.subs, shape.minBlocks);
	prog.cluster = shape.cluster;
	applyJitKnobs(prog);
	std::vector<char> cubin;
	const std::string err = compileFlowKernel(d->core, shape.nb, shape.nbt, shape.subs, shape.cluster, threads * shape.subs, shape.minBlocks, KernelSizes{ L, paddedSites(L), channelsOf(d->core) * paddedSites(L), d->n_frequencies }, generateRpaSource(prog), cubin);
	if (!err.empty()) return fail(PFFRG_ERR_CUDA, "%s", err.c_str());
	if (cubinBytes) *cubinBytes = (int64_t)cubin.size();
	return PFFRG_OK;
}

int pffrg_plan_partition(int core, int nFrequencies, const double *frequencies, int nSites, int64_t rpaTerms, double cutoff, int nRanks, int64_t *bounds)
{
	if (core < 0 || core > 2 || nFrequencies < 2 || !frequencies || nSites < 1 || nRanks < 1 || !bounds) return fail(PFFRG_ERR_ARGUMENT, "bad argument");
	std::vector<int> counts(nFrequencies);
	for (int i = 0; i < nFrequencies; ++i) counts[i] = nodeCount(frequencies, nFrequencies, cutoff, frequencies[i]);
	const std::vector<int64_t> b = planPartition(core, nFrequencies, nSites, (double)rpaTerms, counts, nRanks);
	std::copy(b.begin(), b.end(), bounds);
	return PFFRG_OK;
}

int pffrg_plan_partition_feedback(int core, int nFrequencies, const double *frequencies, int nSites, int64_t rpaTerms, double cutoff, int nRanks,
                                  const int64_t *prevBounds, const double *prevMs, int64_t *bounds)
{
	if (core < 0 || core > 2 || nFrequencies < 2 || !frequencies || nSites < 1 || nRanks < 1 || !bounds || !prevBounds || !prevMs) return fail(PFFRG_ERR_ARGUMENT, "bad argument");
	std::vector<int> counts(nFrequencies);
	for (int i = 0; i < nFrequencies; ++i) counts[i] = nodeCount(frequencies, nFrequencies, cutoff, frequencies[i]);
	const std::vector<int64_t> pb(prevBounds, prevBounds + nRanks + 1);
	const std::vector<double> pt(prevMs, prevMs + nRanks);
	const std::vector<int64_t> b = planPartition(core, nFrequencies, nSites, (double)rpaTerms, counts, nRanks, &pb, &pt);
	std::copy(b.begin(), b.end(), bounds);
	return PFFRG_OK;
}

int pffrg_tri_terms(int region, int32_t *terms, int capacity)
{
	// region 0: pp ladder, 1: ph ladder, 2: chalice, 3: inverse chalice (rows {out, sign, first, second}), 4: RPA (rows {out, sign, mu*4+k, k*4+nu})
	if (region < 0 || region > 5 || (!terms && capacity > 0)) return fail(PFFRG_ERR_ARGUMENT, "bad region or null buffer");
	int n = 0;
	auto emit = [&](const tri::Term &t, int first, int second)
	{
		if (t.exponent & 1) return false; // an imaginary coefficient would mean the algebra is inconsistent
		if (n < capacity) { terms[4 * n] = t.out; terms[4 * n + 1] = (int)t.sign; terms[4 * n + 2] = first; terms[4 * n + 3] = second; }
		++n;
		return true;
	};
	bool consistent = true;
	if (region == 5)
	{
		// egg diagram of the correlation measurement: rows {out channel, 4 * coefficient, a, b}
		for (int c = 0; c < 16; ++c)
			for (int ab = 0; ab < 16; ++ab)
			{
				const int mu = c >> 2, nu = c & 3;
				if ((mu == 3) != (nu == 3)) continue;
				const tri::Term t = tri::egg(mu, nu, ab >> 2, ab & 3);
				if (t.sign == 0.0) continue;
				if (t.exponent & 1) { consistent = false; continue; }
				if (n < capacity) { terms[4 * n] = c; terms[4 * n + 1] = (int)(4.0 * t.sign); terms[4 * n + 2] = ab >> 2; terms[4 * n + 3] = ab & 3; }
				++n;
			}
	}
	else if (region == 4)
	{
		for (int mu = 0; mu < 4; ++mu) for (int k = 0; k < 4; ++k) for (int nu = 0; nu < 4; ++nu) consistent &= emit(tri::rpa(mu, k, nu), 4 * mu + k, 4 * k + nu);
	}
	else
		for (int c1 = 0; c1 < 16; ++c1)
			for (int c2 = 0; c2 < 16; ++c2)
			{
				const int a = c1 >> 2, b = c1 & 3, g = c2 >> 2, dd = c2 & 3;
				const tri::Term t = region == 0 ? tri::ladder<false>(a, b, g, dd) : region == 1 ? tri::ladder<true>(a, b, g, dd) : region == 2 ? tri::chalice(a, b, g, dd) : tri::inverseChalice(a, b, g, dd);
				consistent &= emit(t, c1, c2);
			}
	if (!consistent) return fail(PFFRG_ERR_STATE, "TRI spin algebra produced an imaginary coefficient");
	return n;
}

int pffrg_trigram_tables(const pffrg_desc *d, int resident, int warps, uint16_t *blocks, int blockCapacity, uint32_t *terms, int capacity, int32_t *seg, int segCapacity, int32_t *rounds)
{
	if (!d || d->n_sites < 1 || 16 * d->n_sites > 1024 || !d->overlap_offsets || resident < 1 || warps < 1 || warps > 32 || !rounds) return fail(PFFRG_ERR_ARGUMENT, "bad argument");
	std::vector<unsigned short> b; std::vector<unsigned> t; std::vector<int> s;
	buildTriGramTables(d, d->n_sites, resident, warps, b, t, s);
	*rounds = (int)(b.size() / resident);
	if (blocks) std::copy(b.begin(), b.begin() + std::min<size_t>(b.size(), (size_t)std::max(blockCapacity, 0)), blocks);
	if (terms) std::copy(t.begin(), t.begin() + std::min<size_t>(t.size(), (size_t)std::max(capacity, 0)), terms);
	if (seg) std::copy(s.begin(), s.begin() + std::min<size_t>(s.size(), (size_t)std::max(segCapacity, 0)), seg);
	return (int)t.size();
}

int pffrg_site_order(const pffrg_desc *d, int32_t *order)
{
	if (!d || d->n_sites < 1 || !d->inverted_rid || !d->sites_rid || !d->sites_perm || !d->inverted_perm || !d->overlap_offsets || !d->range_fwd_rid || !d->range_inv_rid || !order) return fail(PFFRG_ERR_ARGUMENT, "bad argument");
	const RelabelledDesc r(d);
	std::copy(r.order.begin(), r.order.end(), order);
	return r.identity ? 0 : 1;
}

int pffrg_gram_tables(const pffrg_desc *d, int rowsPerBlock, int warps, uint32_t *terms, int capacity, int32_t *seg, double *conflictDegree)
{
	if (!d || d->n_sites < 1 || d->n_sites > 256 || !d->overlap_offsets || rowsPerBlock < 1 || warps < 1 || warps > 32 || !seg || (!terms && capacity > 0)) return fail(PFFRG_ERR_ARGUMENT, "bad argument");
	const int L = d->n_sites, Lp = gramGeometry(d->core, L, paddedSites(L)).lp; // (XYZ: three virtual sites per site)
	if (d->core == XYZ && 4 * L > 255) return fail(PFFRG_ERR_UNSUPPORTED, "the XYZ Gram form handles at most 63 representative sites");
	if ((long)rowsPerBlock * (Lp + 1) > (1l << 14) || rowsPerBlock % 8) return fail(PFFRG_ERR_ARGUMENT, "%d rows of %d do not fit the 14 offset bits of a term word (or not a multiple of 8)", rowsPerBlock, Lp + 1);
	std::vector<unsigned> t; std::vector<int> s;
	buildGramTables(d, L, Lp, rowsPerBlock, warps, t, s, conflictDegree);
	std::copy(s.begin(), s.end(), seg);
	std::copy(t.begin(), t.begin() + std::min<size_t>(t.size(), (size_t)std::max(capacity, 0)), terms);
	return (int)t.size();
}

double pffrg_fp64_peak(int device)
{
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); fail(PFFRG_ERR_CUDA, "cudaSetDevice(%d) failed", device); return -1.0; }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return -1.0; }
	double *out = nullptr; cudaEvent_t e0, e1;
	if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) { cudaGetLastError(); return -1.0; }
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = prop.multiProcessorCount * 8, iterations = 20000;
	double best = 0.0;
	for (int rep = 0; rep < 4; ++rep)
	{
		cudaEventRecord(e0);
		fp64PeakKernel<<<blocks, 256>>>(out, iterations, 1.0000001, 0.9999999);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
		const double tflops = 2.0 * 16.0 * iterations * 256.0 * blocks / (ms * 1e-3) / 1e12;
		if (rep > 0 && tflops > best) best = tflops;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
	if (cudaGetLastError() != cudaSuccess) { fail(PFFRG_ERR_CUDA, "FP64 probe failed"); return -1.0; }
	return best;
}

double pffrg_dmma_peak(int device)
{
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); fail(PFFRG_ERR_CUDA, "cudaSetDevice(%d) failed", device); return -1.0; }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return -1.0; }
	double *out = nullptr; cudaEvent_t e0, e1;
	if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) { cudaGetLastError(); return -1.0; }
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = prop.multiProcessorCount * 8, iterations = 4000;
	double best = 0.0;
	for (int rep = 0; rep < 4; ++rep)
	{
		cudaEventRecord(e0);
		dmmaPeakKernel<<<blocks, 256>>>(out, iterations, 1.0000001, 0.9999999);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
		// one m8n8k4 = 256 multiply-adds per warp
		const double tflops = 2.0 * 256.0 * 8.0 * iterations * (256.0 / 32.0) * blocks / (ms * 1e-3) / 1e12;
		if (rep > 0 && tflops > best) best = tflops;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
	if (cudaGetLastError() != cudaSuccess) { fail(PFFRG_ERR_CUDA, "FP64 tensor-core probe failed"); return -1.0; }
	return best;
}

void *pffrg_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); fail(PFFRG_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
	return p;
}
void pffrg_host_free(void *p) { if (p) cudaFreeHost(p); }

int pffrg_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return fail(PFFRG_ERR_ARGUMENT, "null pointer or empty range");
	CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
	return PFFRG_OK;
}
int pffrg_host_unregister(void *p)
{
	if (!p) return fail(PFFRG_ERR_ARGUMENT, "null pointer");
	CUDA_TRY(cudaHostUnregister(p));
	return PFFRG_OK;
}

} // extern "C"
