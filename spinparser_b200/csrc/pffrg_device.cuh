// pffrg_device.cuh -- device-side building blocks of the pf-FRG flow kernels (sm_100a, FP64).
//
// Everything here restates, for the GPU, semantics fixed by the reference (file:line relative to the SpinParser tree):
//   frequency mesh search / lerp          src/FrequencyDiscretization.hpp:253-351
//   self-energy access (odd, clamped)     src/SU2/SU2VertexSingleParticle.hpp:73-87
//   access buffers (sector map + lerps)   src/SU2/SU2VertexTwoParticle.hpp:399-490, src/TRI/TRIVertexTwoParticle.hpp:401-504
// The arithmetic expressions keep the reference's operation order where it decides which mesh cell a frequency falls
// into; summation order elsewhere is free (parity tolerance 1e-10, round-off head-room ~1e-15).
#pragma once

#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#endif

namespace pffrg
{
	enum Core : int { SU2 = 0, XYZ = 1, TRI = 2 };
	enum Channel : int { CH_S = 0, CH_T = 1, CH_U = 2, CH_NONE = 3 };

	__host__ __device__ constexpr int channelsOf(int core) { return core == SU2 ? 2 : (core == XYZ ? 4 : 16); }

	// Device layout of the vertex: v4[row][c / VW][site][c % VW] with VW = 2 for SU2 and XYZ -- the channels of a site are stored in
	// interleaved pairs, so one 16-byte load fetches a pair and a warp reads 32 * 16 contiguous bytes -- and VW = 1 for TRI, whose
	// gathers address single (spin-permuted) channels. A row is RL = C * Lp doubles in both cases.
	__host__ __device__ constexpr int vectorWidth(int core) { return core == TRI ? 1 : 2; }
	__host__ __device__ __forceinline__ int channelOffset(int vw, int c, int Lp) { return (c / vw) * (vw * Lp) + (c % vw); }

	// ---- frequency mesh -------------------------------------------------------------------------------------------
	// first index i in [1, nw) with mesh[i] > w, or nw if none (the linear scans of FrequencyDiscretization.hpp:264-267,
	// 290-293, 338-347 as a binary search)
	__host__ __device__ __forceinline__ int firstGreater(const double *mesh, int nw, double w)
	{
		int lo = 1, hi = nw;
		while (lo < hi)
		{
			int mid = (lo + hi) >> 1;
			if (mesh[mid] > w) hi = mid; else lo = mid + 1;
		}
		return lo;
	}

	// FrequencyDiscretization::interpolateOffset, :326-351
	__host__ __device__ __forceinline__ void interpolateOffset(const double *mesh, int nw, double w, int &lower, int &upper, double &bias)
	{
		if (w <= mesh[0]) { lower = 0; upper = 0; bias = 0.0; return; }
		int i = firstGreater(mesh, nw, w);
		if (i >= nw) { lower = nw - 1; upper = nw - 1; bias = 0.0; return; }
		upper = i; lower = i - 1;
		bias = (w - mesh[lower]) / (mesh[upper] - mesh[lower]);
	}

	// FrequencyDiscretization::offset, :306-316: first i >= 1 with mesh[i] >= w
	__host__ __device__ __forceinline__ int exactOffset(const double *mesh, int nw, double w)
	{
		if (w <= mesh[0]) return 0;
		int lo = 1, hi = nw;
		while (lo < hi)
		{
			int mid = (lo + hi) >> 1;
			if (mesh[mid] >= w) hi = mid; else lo = mid + 1;
		}
		return lo < nw ? lo : nw - 1;
	}

	// mesh value at a signed index (negative half mirrored: value(-i-1) = -mesh[i], :187-192)
	__host__ __device__ __forceinline__ double meshValue(const double *mesh, int index) { return index >= 0 ? mesh[index] : -mesh[-index - 1]; }

	// FrequencyDiscretization::lesser / greater, :253-296, signed-index form, including the quirk for |w| <= mesh[0]
	__host__ __device__ __forceinline__ int meshGreaterPos(const double *mesh, int nw, double w)
	{
		if (w <= mesh[0]) return 0;
		int i = firstGreater(mesh, nw, w);
		return i < nw ? i : nw - 1;
	}
	__host__ __device__ __forceinline__ int meshLesserPos(const double *mesh, int nw, double w)
	{
		if (w <= mesh[0]) return 0;
		int i = firstGreater(mesh, nw, w);
		return i < nw ? i - 1 : nw - 1;
	}
	__host__ __device__ __forceinline__ int meshLesser(const double *mesh, int nw, double w)
	{
		return w < 0 ? -(meshGreaterPos(mesh, nw, -w) + 1) : meshLesserPos(mesh, nw, w);
	}
	__host__ __device__ __forceinline__ int meshGreater(const double *mesh, int nw, double w)
	{
		return w < 0 ? -(meshLesserPos(mesh, nw, -w) + 1) : meshGreaterPos(mesh, nw, w);
	}

	// {SU2,XYZ,TRI}VertexSingleParticle::getValue, src/SU2/SU2VertexSingleParticle.hpp:73-87
	__host__ __device__ __forceinline__ double selfEnergy(const double *mesh, int nw, const double *v2, double w)
	{
		double sign = 1.0;
		if (w < 0) { w = -w; sign = -1.0; }
		int lo, up; double bias;
		interpolateOffset(mesh, nw, w, lo, up, bias);
		return sign * ((1 - bias) * v2[lo] + bias * v2[up]);
	}

	// ---- quadrature node lists ---------------------------------------------------------------------------------------
	// Number of kernel evaluations of one channel with transfer frequency x at cutoff L: the conventional single-scale
	// terms (src/SU2/SU2FrgCore.cpp:351-371) plus the nodes of the three Katanin segments (:378-392) as enumerated by
	// ImplicitIntegrator (src/lib/Integrator.hpp:138-287).
	__host__ __device__ inline int nodeCount(const double *mesh, int nw, double cutoff, double x)
	{
		int n = 1;
		if (x > 2.0 * cutoff) n += 1;
		if (-(x + cutoff) > -mesh[nw - 1])
		{
			int umax = meshLesser(mesh, nw, -(x + cutoff));
			n += (umax != -nw) ? (umax + nw) + 2 : 2;
		}
		if (x - cutoff > cutoff)
		{
			int umin = meshGreater(mesh, nw, cutoff - x), umax = meshLesser(mesh, nw, -cutoff);
			n += (umax >= umin) ? (umax - umin) + 3 : 2;
		}
		if (cutoff < mesh[nw - 1])
		{
			int umin = meshGreater(mesh, nw, cutoff);
			n += (umin != nw - 1) ? (nw - 1 - umin) + 2 : 2;
		}
		return n;
	}

	// ---- access buffers --------------------------------------------------------------------------------------------------
	// Four interpolation supports of one vertex access: row index (su*Nw + t), weight, and the symmetry flags that decide
	// signs and the site/spin maps at gather time.
	// 80 bytes = an odd number of 16-byte words: consecutive threads write consecutive entries without bank conflicts.
	struct __align__(16) AccessBuffer
	{
		double w[4];     // interpolation weights; channels that are odd under the s<->u frequency exchange use -w[k] where
		                 // support k reads the mirrored entry (oddWeight)
		int row[4];      // row index su*Nw + t
		int flags;       // bit0: site/pair exchange, bit1: TRI zeta_mu*zeta_nu factor, bit(4+k): support k reads the s<->u mirrored entry
		int pad[7];
	};
	constexpr int AB_EXCHANGE = 1, AB_TZ = 2;
	__host__ __device__ __forceinline__ int abSwapped(int flags, int k) { return (flags >> (4 + k)) & 1; }
	// weight of support k for a channel that is odd under s<->u: sign bit flipped iff the support is mirrored
	__device__ __forceinline__ double oddWeight(double w, int flags, int k)
	{
		return __hiloint2double(__double2hiint(w) ^ (int)(((unsigned)flags >> (4 + k)) << 31), __double2loint(w));
	}

	__host__ __device__ __forceinline__ int rowIndex(int nw, int so, int to, int uo, int k, int &flags)
	{
		if (so < uo) { flags |= 1 << (4 + k); return (uo * (uo + 1) / 2 + so) * nw + to; }
		return (so * (so + 1) / 2 + uo) * nw + to;
	}

	// generateAccessBuffer(s, t, u, channel): SU2VertexTwoParticle.hpp:399-490 (XYZ identical), TRIVertexTwoParticle.hpp:401-504
	template <int CORE>
	__host__ __device__ inline void makeAccessBuffer(const double *mesh, int nw, double s, double t, double u, int channel, AccessBuffer &ab)
	{
		int flags = 0;
		if (CORE == TRI)
		{
			if (s < 0) { s = -s; flags ^= AB_EXCHANGE; }
			if (t < 0) { t = -t; flags ^= AB_TZ; }
			if (u < 0) { u = -u; flags ^= AB_EXCHANGE; flags ^= AB_TZ; }
		}
		else
		{
			if (s < 0 && u < 0) { s = -s; u = -u; }
			else
			{
				if (s < 0) { s = -s; flags |= AB_EXCHANGE; }
				else if (u < 0) { u = -u; flags |= AB_EXCHANGE; }
			}
			if (t < 0) t = -t;
		}
		int l1, u1, l2, u2; double b1, b2;
		if (channel == CH_S)
		{
			int es = exactOffset(mesh, nw, s);
			interpolateOffset(mesh, nw, t, l1, u1, b1);
			interpolateOffset(mesh, nw, u, l2, u2, b2);
			ab.w[0] = (1 - b2) * (1 - b1); ab.row[0] = rowIndex(nw, es, l1, l2, 0, flags);
			ab.w[1] = (1 - b2) * b1;       ab.row[1] = rowIndex(nw, es, u1, l2, 1, flags);
			ab.w[2] = b2 * (1 - b1);       ab.row[2] = rowIndex(nw, es, l1, u2, 2, flags);
			ab.w[3] = b2 * b1;             ab.row[3] = rowIndex(nw, es, u1, u2, 3, flags);
		}
		else if (channel == CH_T)
		{
			int et = exactOffset(mesh, nw, t);
			interpolateOffset(mesh, nw, s, l1, u1, b1);
			interpolateOffset(mesh, nw, u, l2, u2, b2);
			ab.w[0] = (1 - b2) * (1 - b1); ab.row[0] = rowIndex(nw, l1, et, l2, 0, flags);
			ab.w[1] = (1 - b2) * b1;       ab.row[1] = rowIndex(nw, u1, et, l2, 1, flags);
			ab.w[2] = b2 * (1 - b1);       ab.row[2] = rowIndex(nw, l1, et, u2, 2, flags);
			ab.w[3] = b2 * b1;             ab.row[3] = rowIndex(nw, u1, et, u2, 3, flags);
		}
		else
		{
			int eu = exactOffset(mesh, nw, u);
			interpolateOffset(mesh, nw, s, l1, u1, b1);
			interpolateOffset(mesh, nw, t, l2, u2, b2);
			ab.w[0] = (1 - b2) * (1 - b1); ab.row[0] = rowIndex(nw, l1, l2, eu, 0, flags);
			ab.w[1] = (1 - b2) * b1;       ab.row[1] = rowIndex(nw, u1, l2, eu, 1, flags);
			ab.w[2] = b2 * (1 - b1);       ab.row[2] = rowIndex(nw, l1, u2, eu, 2, flags);
			ab.w[3] = b2 * b1;             ab.row[3] = rowIndex(nw, u1, u2, eu, 3, flags);
		}
		ab.flags = flags;
	}

	// ---- access buffers from shared interpolation records ------------------------------------------------------------
	// The 4 (s, u channel) or 8 (t channel) access buffers of one quadrature node use only FOUR distinct interpolated
	// frequencies q0..q3 (up to sign) next to the item's own on-mesh frequency, e.g. for the s channel
	// (src/SU2/SU2FrgCore.cpp:202-206) q = {w1+w', w2+w', w1p+w', w2p+w'} and the buffers read (t,u) = (-q0,-q1), (q2,-q3),
	// (q1,q0), (-q3,q2). So the mesh searches are done once per (node, q) -- LerpRecord -- and the buffers are assembled
	// from two records each without any search. (-a-b == -(a+b) and a-b == -(b-a) hold exactly in IEEE arithmetic, so the
	// interpolation arguments are bit-identical to the reference's.)
	struct __align__(16) LerpRecord
	{
		double bias;
		short lower, upper;
		short neg, pos; // q < 0, q > 0
	};

	// Bucket index over the frequency mesh: start[k] = firstGreater(lower edge of bucket k), buckets being runs of equal
	// leading bits of the IEEE representation (monotone for positive doubles). A search touches start[k], start[k+1] and,
	// for the usual meshes, at most one mesh value instead of log2(Nw) dependent shared-memory loads.
	struct MeshIndex
	{
		const unsigned short *start; // [nKeys + 1]
		int shift;                   // key = (bits >> shift) - keyBase
		int keyBase;
		int nKeys;
	};

	__device__ __forceinline__ int firstGreaterIndexed(const double *mesh, int nw, const MeshIndex &ix, double w)
	{
		const long long key = (long long)(__double_as_longlong(w) >> ix.shift) - ix.keyBase;
		if (key >= ix.nKeys) return nw;
		int a = ix.start[key], b = ix.start[key + 1];
		while (a < b)
		{
			const int mid = (a + b) >> 1;
			if (mesh[mid] > w) b = mid; else a = mid + 1;
		}
		return a;
	}

	// interpolateOffset(|q|) through the bucket index (same results as interpolateOffset above)
	__device__ __forceinline__ void makeLerpRecord(const double *mesh, int nw, const MeshIndex &ix, double q, LerpRecord &r)
	{
		r.neg = q < 0 ? 1 : 0; r.pos = q > 0 ? 1 : 0;
		const double w = fabs(q);
		if (w <= mesh[0]) { r.lower = 0; r.upper = 0; r.bias = 0.0; return; }
		const int i = firstGreaterIndexed(mesh, nw, ix, w);
		if (i >= nw) { r.lower = (short)(nw - 1); r.upper = (short)(nw - 1); r.bias = 0.0; return; }
		r.upper = (short)i; r.lower = (short)(i - 1);
		r.bias = (w - mesh[i - 1]) / (mesh[i] - mesh[i - 1]);
	}

	// interpolated frequency q (0..3) of channel ch at integration frequency wp; f = {w1p, w1, w2p, w2} of the item
	__device__ __forceinline__ double nodeQuantity(int ch, int q, double w1p, double w1, double w2p, double w2, double wp)
	{
		if (ch == CH_S) return q == 0 ? w1 + wp : (q == 1 ? w2 + wp : (q == 2 ? w1p + wp : w2p + wp));
		if (ch == CH_U) return q == 0 ? w1 + wp : (q == 1 ? wp - w2p : (q == 2 ? w1p + wp : w2 - wp));
		return q == 0 ? w1 - wp : (q == 1 ? w1p + wp : (q == 2 ? w2p - wp : w2 + wp));
	}

	// (channel, buffer) -> which records form the two interpolated arguments: q1 | flip1 << 2 | q2 << 3 | flip2 << 5 | exactNegative << 6.
	// s channel: (first, second) = (t, u); u channel and the t channel's site-0 buffers 4..7: (s, t); t channel buffers 0..3: (s, u).
	__host__ __device__ constexpr unsigned long long packRecipes(const int (&t)[8]) { unsigned long long p = 0; for (int i = 0; i < 8; ++i) p |= (unsigned long long)t[i] << (7 * i); return p; }
	__device__ __forceinline__ int bufferRecipe(int ch, int b)
	{
		constexpr int F = 4, G = 32, N = 64;
		constexpr int recS[8] = { 0 | F | (1 << 3) | G, 2 | (3 << 3) | G, 1 | (0 << 3), 3 | F | (2 << 3), 0, 0, 0, 0 };
		constexpr int recU[8] = { 0 | (1 << 3), 2 | (3 << 3), 1 | F | (0 << 3) | G, 3 | (2 << 3), 0, 0, 0, 0 };
		constexpr int recT[8] = { 0 | (1 << 3), 2 | (3 << 3) | G, 1 | (0 << 3), 3 | (2 << 3) | G,
		                          2 | (3 << 3) | G, 0 | (1 << 3) | G | N, 3 | (2 << 3) | G, 1 | (0 << 3) | G | N };
		constexpr unsigned long long pS = packRecipes(recS), pU = packRecipes(recU), pT = packRecipes(recT);
		const unsigned long long p = ch == CH_S ? pS : (ch == CH_U ? pU : pT);
		return (int)((p >> (7 * b)) & 127u);
	}

	// generateAccessBuffer (SU2VertexTwoParticle.hpp:399-490, TRIVertexTwoParticle.hpp:401-504) from two interpolation records;
	// `exactIndex` is the mesh index of the on-mesh argument (offset() of a mesh value is its own index, FrequencyDiscretization.hpp:306-316)
	template <int CORE>
	__device__ __forceinline__ void assembleFromRecords(int nw, int ch, int b, int exactIndex, int recipe, const LerpRecord &r1, const LerpRecord &r2, AccessBuffer &ab);
	template <int CORE>
	__device__ __forceinline__ void assembleAccessBuffer(int nw, int ch, int b, int exactIndex, const LerpRecord *rec /* [4] */, AccessBuffer &ab)
	{
		const int recipe = bufferRecipe(ch, b);
		assembleFromRecords<CORE>(nw, ch, b, exactIndex, recipe, rec[recipe & 3], rec[(recipe >> 3) & 3], ab);
	}
	template <int CORE>
	__device__ __forceinline__ void assembleFromRecords(int nw, int ch, int b, int exactIndex, int recipe, const LerpRecord &r1, const LerpRecord &r2, AccessBuffer &ab)
	{
		const bool neg1 = (recipe & 4) ? r1.pos != 0 : r1.neg != 0, neg2 = (recipe & 32) ? r2.pos != 0 : r2.neg != 0;
		const bool local = ch == CH_T && b >= 4;
		bool sNeg, tNeg, uNeg;
		if (ch == CH_S) { sNeg = false; tNeg = neg1; uNeg = neg2; }
		else if (ch == CH_T && !local) { sNeg = neg1; tNeg = false; uNeg = neg2; }
		else { sNeg = neg1; tNeg = neg2; uNeg = (recipe & 64) != 0; }
		int flags = 0;
		if (CORE == TRI)
		{
			if (sNeg) flags ^= AB_EXCHANGE;
			if (tNeg) flags ^= AB_TZ;
			if (uNeg) { flags ^= AB_EXCHANGE; flags ^= AB_TZ; }
		}
		else if (sNeg != uNeg) flags |= AB_EXCHANGE;
		const int l1 = r1.lower, u1 = r1.upper, l2 = r2.lower, u2 = r2.upper;
		const double b1 = r1.bias, b2 = r2.bias;
		ab.w[0] = (1 - b2) * (1 - b1); ab.w[1] = (1 - b2) * b1; ab.w[2] = b2 * (1 - b1); ab.w[3] = b2 * b1;
		if (ch == CH_S)
		{
			ab.row[0] = rowIndex(nw, exactIndex, l1, l2, 0, flags); ab.row[1] = rowIndex(nw, exactIndex, u1, l2, 1, flags);
			ab.row[2] = rowIndex(nw, exactIndex, l1, u2, 2, flags); ab.row[3] = rowIndex(nw, exactIndex, u1, u2, 3, flags);
		}
		else if (ch == CH_T && !local)
		{
			ab.row[0] = rowIndex(nw, l1, exactIndex, l2, 0, flags); ab.row[1] = rowIndex(nw, u1, exactIndex, l2, 1, flags);
			ab.row[2] = rowIndex(nw, l1, exactIndex, u2, 2, flags); ab.row[3] = rowIndex(nw, u1, exactIndex, u2, 3, flags);
		}
		else
		{
			ab.row[0] = rowIndex(nw, l1, l2, exactIndex, 0, flags); ab.row[1] = rowIndex(nw, u1, l2, exactIndex, 1, flags);
			ab.row[2] = rowIndex(nw, l1, u2, exactIndex, 2, flags); ab.row[3] = rowIndex(nw, u1, u2, exactIndex, 3, flags);
		}
		ab.flags = flags;
	}

	// TRIVertexTwoParticle::_zeta, src/TRI/TRIVertexTwoParticle.hpp:674-677
	__host__ __device__ __forceinline__ double zeta(int c) { return c <= 2 ? -1.0 : 1.0; }

	// sign of support k for OUTPUT channel c (SU2VertexTwoParticle.hpp:377,616-632; XYZ :393,:640-655; TRI :412-443,:643-666)
	template <int CORE>
	__host__ __device__ __forceinline__ double supportSign(int flags, int k, int c)
	{
		if (CORE == SU2) return (c == 1 && abSwapped(flags, k)) ? -1.0 : 1.0;
		if (CORE == XYZ) return (c == 3 && abSwapped(flags, k)) ? -1.0 : 1.0;
		int mu = c >> 2, nu = c & 3;
		double sgn = 1.0;
		if (flags & AB_TZ) sgn *= zeta(mu) * zeta(nu);
		if (abSwapped(flags, k)) sgn *= (flags & AB_EXCHANGE) ? -zeta(mu) : -zeta(nu);
		return sgn;
	}

	// stored channel read by output channel c at a site with spin permutation perm (packed 2 bits per component)
	// XYZVertexTwoParticle.hpp:401-404, TRIVertexTwoParticle.hpp:378-382
	template <int CORE>
	__host__ __device__ __forceinline__ int storedChannel(int flags, int c, int perm)
	{
		if (CORE == SU2) return c;
		if (CORE == XYZ) return c < 3 ? ((perm >> (2 * c)) & 3) : 3;
		int mu = c >> 2, nu = c & 3;
		int m = (flags & AB_EXCHANGE) ? nu : mu, n = (flags & AB_EXCHANGE) ? mu : nu;
		if (m < 3) m = (perm >> (2 * m)) & 3;
		if (n < 3) n = (perm >> (2 * n)) & 3;
		return 4 * m + n;
	}
	constexpr int PERM_IDENTITY = 0 | (1 << 2) | (2 << 4);

	// ---- spin algebra of the TRI core ------------------------------------------------------------------------------------
	// The TRI vertex is Gamma = sum_{mu,nu} Gamma^{mu nu} Theta^{mu nu}, Theta^{mu nu} = c_{mu nu} sigma^mu (x) sigma^nu with
	// sigma^{0,1,2} the Pauli matrices, sigma^3 (the "density" component d) the identity, and c = i for the mixed spin-density
	// components, 1 otherwise (which keeps every Gamma^{mu nu} real for time-reversal invariant models). Every bilinear term of
	// the flow equations is the coefficient of Theta^{mu nu} in a product of two such operators:
	//   pp ladder        + [ (sigma^g sigma^a) (x) (sigma^d sigma^b) ]         A^{ab} B^{gd}   (src/TRI/TRIFrgCore.cpp:198-711)
	//   ph ladder        - [ (sigma^g sigma^a) (x) (sigma^b sigma^d) ]         A^{ab} B^{gd}   (:2389-2902)
	//   chalice          - [ sigma^a (x) (sigma^d sigma^b sigma^g) ]           A^{ab} B0^{gd}  (:1298-1811), B0 = site-0 value
	//   inverse chalice  - [ (sigma^d sigma^a sigma^g) (x) sigma^b ]           A^{ab} B0^{gd}  (:1855-2368)
	//   RPA              + 2 c_{mu k} c_{k nu} / c_{mu nu}                      A^{mu k}[r1] B^{k nu}[r2]  (:739-1252)
	// Products of Pauli matrices only generate phases i^n, so the whole algebra is integer arithmetic modulo 4 and can be
	// evaluated at compile time. tests/test_tri_tables.py checks the resulting tables term by term against the reference file.
	namespace tri
	{
		__host__ __device__ constexpr int enc(int a) { return (a + 1) & 3; }                     // d -> 0, x -> 1, y -> 2, z -> 3
		// sigma^a sigma^b = i^mulPhase(a,b) sigma^mulIndex(a,b)
		__host__ __device__ constexpr int mulIndex(int a, int b) { return ((enc(a) ^ enc(b)) + 3) & 3; }
		__host__ __device__ constexpr int mulPhase(int a, int b) { return (a == 3 || b == 3 || a == b) ? 0 : (((b - a + 3) % 3 == 1) ? 1 : 3); }
		__host__ __device__ constexpr int mixed(int a, int b) { return ((a == 3) != (b == 3)) ? 1 : 0; } // c_{ab} = i^mixed(a,b)
		// i^e for e even -> +-1 (odd exponents do not occur; the host-side table export asserts that)
		__host__ __device__ constexpr double phaseSign(int e) { return ((e & 3) == 0) ? 1.0 : -1.0; }

		struct Term { int out; int exponent; double sign; };

		// ladder term A^{ab} B^{gd} -> Theta^{mu nu}; PH = particle-hole (u channel), else particle-particle (s channel)
		template <bool PH>
		__host__ __device__ constexpr Term ladder(int a, int b, int g, int d)
		{
			const int mu = mulIndex(g, a), nu = mulIndex(b, d);
			const int e = mixed(a, b) + mixed(g, d) + mulPhase(g, a) + (PH ? mulPhase(b, d) : mulPhase(d, b)) + 4 - mixed(mu, nu);
			return { 4 * mu + nu, e & 3, PH ? -phaseSign(e) : phaseSign(e) };
		}
		// chalice term A^{ab} B0^{gd} -> Theta^{a nu}, nu from sigma^d sigma^b sigma^g
		__host__ __device__ constexpr Term chalice(int a, int b, int g, int d)
		{
			const int m1 = mulIndex(d, b), nu = mulIndex(m1, g);
			const int e = mixed(a, b) + mixed(g, d) + mulPhase(d, b) + mulPhase(m1, g) + 4 - mixed(a, nu);
			return { 4 * a + nu, e & 3, -phaseSign(e) };
		}
		// inverse chalice term A^{ab} B0^{gd} -> Theta^{mu b}, mu from sigma^d sigma^a sigma^g
		__host__ __device__ constexpr Term inverseChalice(int a, int b, int g, int d)
		{
			const int m1 = mulIndex(d, a), mu = mulIndex(m1, g);
			const int e = mixed(a, b) + mixed(g, d) + mulPhase(d, a) + mulPhase(m1, g) + 4 - mixed(mu, b);
			return { 4 * mu + b, e & 3, -phaseSign(e) };
		}
		// RPA term A^{mu k} B^{k nu} -> Theta^{mu nu} (the factor 2 = tr(sigma sigma) is applied by the caller)
		__host__ __device__ constexpr Term rpa(int mu, int k, int nu)
		{
			const int e = mixed(mu, k) + mixed(k, nu) + 4 - mixed(mu, nu);
			return { 4 * mu + nu, e & 3, phaseSign(e) };
		}
	}
}
