// pffrg_measure.cuh -- K5: static spin-spin correlations chi^{mu nu}_{0j}(Lambda) from the flowing vertex, on the device.
//
// Restates {SU2,XYZ,TRI}MeasurementCorrelation::_calculateCorrelation
// (src/SU2/SU2MeasurementCorrelation.cpp:77-177, src/XYZ/XYZMeasurementCorrelation.cpp:98-225, src/TRI/TRIMeasurementCorrelation.cpp:147-330)
// at external frequency nu = 0:
//   chi_c[r] = int dw { delta_{r0} a_c P(w) + int dw' N(w,w') ( -b_c Gamma_c(r; w+w', 0, w-w') + delta_{r0} e_c[Gamma(0; w+w', w-w', 0)] ) }
// with P(w) = 1/G(w)^2, N = P(w) P(w') / 4 pi^2, G(w) = w + Sigma(w), both integrals over the two outer segments of the sharp
// cutoff (ImplicitIntegrator, src/lib/Integrator.hpp:138-287 -- the same node lists as the Katanin integrals of the flow),
// Gamma evaluated with the 8-support trilinear interpolation of generateAccessBuffer(s,t,u) (SU2VertexTwoParticle.hpp:500-557).
// In the reference this is ONE single-threaded work item per cutoff step; here one CTA per outer node, thread groups over
// inner nodes, threads over representative sites; a second kernel adds the outer nodes in a fixed order.
#pragma once

#include "pffrg_kernels.cuh"

namespace pffrg
{
	struct AccessBuffer8
	{
		double w[8];
		int row[8];
		int flags; // as AccessBuffer: bit0 exchange, bit1 TRI zeta factor, bit(4+n) support n reads the s<->u mirrored entry
	};

	// generateAccessBuffer(s, t, u): SU2VertexTwoParticle.hpp:500-557 (XYZ identical), TRIVertexTwoParticle.hpp:514-577
	template <int CORE>
	__device__ inline void makeAccessBuffer8(const double *mesh, int nw, double s, double t, double u, AccessBuffer8 &ab)
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
		int ls, us, lt, ut, lu, uu; double bs, bt, bu;
		interpolateOffset(mesh, nw, s, ls, us, bs);
		interpolateOffset(mesh, nw, t, lt, ut, bt);
		interpolateOffset(mesh, nw, u, lu, uu, bu);
		#pragma unroll
		for (int n = 0; n < 8; ++n)
		{
			ab.w[n] = ((n & 2) ? bt : 1 - bt) * ((n & 1) ? bs : 1 - bs) * ((n & 4) ? bu : 1 - bu);
			ab.row[n] = rowIndex(nw, (n & 1) ? us : ls, (n & 2) ? ut : lt, (n & 4) ? uu : lu, n, flags);
		}
		ab.flags = flags;
	}

	// getValueSuperbundle / getValueLocal with an 8-support buffer at one site; `local` selects the TRI index order of getValueLocal
	template <int CORE>
	__device__ inline void gatherSite8(const Problem &P, const double *__restrict__ v4, const AccessBuffer8 &ab, int siteFwd, int siteInv, int permFwd, int permInv, bool local, double (&out)[channelsOf(CORE)])
	{
		constexpr int C = channelsOf(CORE);
		const bool exchange = ab.flags & AB_EXCHANGE;
		const int site = exchange ? siteInv : siteFwd;
		const int perm = exchange ? permInv : permFwd;
		#pragma unroll
		for (int c = 0; c < C; ++c)
		{
			const int sc = storedChannel<CORE>(ab.flags, c, perm);
			// TRI: getValueLocal swaps the spin indices under pair exchange BEFORE indexing the sign table (TRIVertexTwoParticle.hpp:349-353)
			const int cs = (CORE == TRI && local && exchange) ? (4 * (c & 3) + (c >> 2)) : c;
			double v = 0.0;
			#pragma unroll
			for (int n = 0; n < 8; ++n) v += supportSign<CORE>(ab.flags, n, cs) * ab.w[n] * __ldg(v4 + (size_t)ab.row[n] * P.RL + channelOffset(vectorWidth(CORE), sc, P.Lp) + site * vectorWidth(CORE));
			out[c] = v;
		}
	}

	namespace tri
	{
		// egg diagram of the TRI correlator: e^{mu nu} += coefficient * v^{ab}, coefficient = (n/4) c_{ab} tr(sigma^mu sigma^b sigma^nu sigma^a),
		// n = 4 for the density-density component, 1 for spin-spin (src/TRI/TRIMeasurementCorrelation.cpp:204-243). `exponent` odd or
		// sign == 0 means no contribution.
		__host__ __device__ constexpr Term egg(int mu, int nu, int a, int b)
		{
			if (mulIndex(mu, b) != mulIndex(nu, a)) return { 4 * mu + nu, 0, 0.0 };
			const int e = mulPhase(mu, b) + mulPhase(nu, a) + mixed(a, b);
			return { 4 * mu + nu, e & 3, ((mu == 3 && nu == 3) ? 2.0 : 0.5) * phaseSign(e) };
		}
	}

	// per-channel constants: a_c (free bubble), b_c (dumbbell), and the egg combination of the site-0 values
	template <int CORE>
	__device__ inline void correlationCoefficients(double spin, double (&a)[channelsOf(CORE)], double (&b)[channelsOf(CORE)])
	{
		constexpr double PI = 3.14159265358979323846;
		if (CORE == SU2) { a[0] = spin / (2.0 * PI); a[1] = 2.0 * spin / PI; b[0] = spin * spin; b[1] = 16.0 * spin * spin; }
		else if (CORE == XYZ) { a[0] = a[1] = a[2] = 1.0 / (4.0 * PI); a[3] = 1.0 / PI; b[0] = b[1] = b[2] = 1.0; b[3] = 4.0; }
		else
		{
			#pragma unroll
			for (int c = 0; c < channelsOf(CORE); ++c)
			{
				const int mu = c >> 2, nu = c & 3;
				const bool dd = mu == 3 && nu == 3, ss = mu < 3 && nu < 3;
				a[c] = dd ? 2.0 / (2.0 * PI) : ((ss && mu == nu) ? 0.5 / (2.0 * PI) : 0.0);
				b[c] = dd ? 4.0 : (ss ? 1.0 : 0.0);
			}
		}
	}
	template <int CORE>
	__device__ inline void eggTerms(double spin, const double (&v)[channelsOf(CORE)], double (&e)[channelsOf(CORE)])
	{
		if (CORE == SU2) { e[0] = spin * (-v[0] / 4.0 + v[1]); e[1] = spin * (3.0 * v[0] + 4.0 * v[1]); }
		else if (CORE == XYZ)
		{
			e[0] = 0.5 * (v[0] - v[1] - v[2] + v[3]);
			e[1] = 0.5 * (-v[0] + v[1] - v[2] + v[3]);
			e[2] = 0.5 * (-v[0] - v[1] + v[2] + v[3]);
			e[3] = 2.0 * (v[0] + v[1] + v[2] + v[3]);
		}
		else
		{
			#pragma unroll
			for (int c = 0; c < channelsOf(CORE); ++c)
			{
				const int mu = c >> 2, nu = c & 3;
				double sum = 0.0;
				if ((mu == 3) == (nu == 3))
				{
					#pragma unroll
					for (int ab = 0; ab < 16; ++ab)
					{
						const tri::Term t = tri::egg(mu, nu, ab >> 2, ab & 3);
						if (t.sign != 0.0 && !(t.exponent & 1)) sum += t.sign * v[ab];
					}
				}
				e[c] = sum;
			}
		}
	}

	// quadrature nodes of int_Lambda (external frequency x) without the conventional terms: integration frequency and trapezoid
	// weight, same enumeration as nodeTableKernel
	__device__ inline int enumerateKataninNodes(const double *mesh, int nw, double cutoff, double x, double *nodeW, double *nodeT)
	{
		int n = 0;
		auto MV = [&](int i) { return meshValue(mesh, i); };
		auto emit = [&](double w, double weight) { nodeW[n] = w; nodeT[n] = weight; ++n; };
		if (-(x + cutoff) > -mesh[nw - 1])
		{
			const double max = -(x + cutoff);
			int umin = -nw;
			const int umax = meshLesser(mesh, nw, max);
			if (umin != umax)
			{
				emit(MV(umin), 0.5 * (MV(umin + 1) - MV(umin)));
				while (++umin != umax) emit(MV(umin), 0.5 * (MV(umin + 1) - MV(umin - 1)));
				emit(MV(umin), 0.5 * (max - MV(umin - 1)));
				emit(max, 0.5 * (max - MV(umin)));
			}
			else { emit(max, 0.5 * (max - MV(umin))); emit(MV(umin), 0.5 * (max - MV(umin))); }
		}
		if (x - cutoff > cutoff)
		{
			const double min = cutoff - x, max = -cutoff;
			int umin = meshGreater(mesh, nw, min);
			const int umax = meshLesser(mesh, nw, max);
			if (umax >= umin)
			{
				emit(min, 0.5 * (MV(umin) - min));
				if (umax != umin)
				{
					emit(MV(umin), 0.5 * (MV(umin + 1) - min));
					while (++umin != umax) emit(MV(umin), 0.5 * (MV(umin + 1) - MV(umin - 1)));
					emit(MV(umin), 0.5 * (max - MV(umin - 1)));
				}
				else emit(MV(umin), 0.5 * (max - min));
				emit(max, 0.5 * (max - MV(umin)));
			}
			else { emit(max, 0.5 * (max - min)); emit(min, 0.5 * (max - min)); }
		}
		if (cutoff < mesh[nw - 1])
		{
			const double min = cutoff;
			const int max = nw - 1;
			int umin = meshGreater(mesh, nw, min);
			if (umin != max)
			{
				emit(min, 0.5 * (MV(umin) - min));
				emit(MV(umin), 0.5 * (MV(umin + 1) - min));
				while (++umin != max) emit(MV(umin), 0.5 * (MV(umin + 1) - MV(umin - 1)));
				emit(MV(umin), 0.5 * (MV(umin) - MV(umin - 1)));
			}
			else { emit(min, 0.5 * (MV(umin) - min)); emit(MV(umin), 0.5 * (MV(umin) - min)); }
		}
		return n;
	}

	// partial[i][c][r] = outer weight_i * ( delta_{r0} a_c P(w_i) + sum_k weight_k N(w_i, w_k) ( -b_c Gamma_c(r) + delta_{r0} e_c ) ); count[0] = number of outer nodes
	template <int CORE>
	__global__ void __launch_bounds__(256) correlationKernel(Problem P, const double *__restrict__ v4, const double *__restrict__ v2, const double *cutoffPtr, int nodeStride, double *__restrict__ partial, int *count)
	{
		constexpr int C = channelsOf(CORE);
		constexpr double PI = 3.14159265358979323846;
		extern __shared__ double smem[];
		const int nw = P.nw, L = P.L;
		double *mesh = smem, *sv2 = smem + nw, *nodeW = smem + 2 * nw, *nodeT = nodeW + nodeStride, *red = nodeT + nodeStride; // red[groups][C][L]
		__shared__ int nNodes;
		for (int i = threadIdx.x; i < nw; i += blockDim.x) { mesh[i] = P.mesh[i]; sv2[i] = v2[i]; }
		__syncthreads();
		const double cutoff = *cutoffPtr, nu = 0.0;
		if (threadIdx.x == 0)
		{
			nNodes = enumerateKataninNodes(mesh, nw, cutoff, nu, nodeW, nodeT);
			if (blockIdx.x == 0) count[0] = nNodes;
		}
		__syncthreads();
		const int n = nNodes;
		if ((int)blockIdx.x >= n) return;
		auto G = [&](double w) { return w + selfEnergy(mesh, nw, sv2, w); };
		const double w = nodeW[blockIdx.x], outerWeight = nodeT[blockIdx.x];
		const double pw = 1.0 / (G(w) * G(w + nu));

		const int groups = blockDim.x / L, g = threadIdx.x / L, j = threadIdx.x - g * L;
		const bool worker = g < groups;
		int siteFwd = 0, siteInv = 0, permFwd = PERM_IDENTITY, permInv = PERM_IDENTITY;
		if (worker) { siteFwd = P.sites_rid[j]; siteInv = P.inv_rid[j]; permFwd = P.sites_perm[j]; permInv = P.inv_perm[j]; }
		double a[C], b[C], acc[C];
		correlationCoefficients<CORE>(P.spin, a, b);
		#pragma unroll
		for (int c = 0; c < C; ++c) acc[c] = 0.0;
		if (worker)
		{
			for (int k = g; k < n; k += groups)
			{
				const double wk = nodeW[k];
				AccessBuffer8 ab;
				makeAccessBuffer8<CORE>(mesh, nw, w + wk + nu, nu, w - wk, ab);
				double stack[C];
				gatherSite8<CORE>(P, v4, ab, siteFwd, siteInv, permFwd, permInv, false, stack);
				const double norm = nodeT[k] * pw / (G(wk) * G(wk + nu) * (4.0 * PI * PI));
				#pragma unroll
				for (int c = 0; c < C; ++c) acc[c] -= norm * b[c] * stack[c];
				if (j == 0)
				{
					// egg diagram: site-0 vertex at (s,t,u) = (w + w' + nu, w - w', nu)
					makeAccessBuffer8<CORE>(mesh, nw, w + wk + nu, w - wk, nu, ab);
					double v[C], e[C];
					gatherSite8<CORE>(P, v4, ab, 0, 0, PERM_IDENTITY, PERM_IDENTITY, true, v);
					eggTerms<CORE>(P.spin, v, e);
					#pragma unroll
					for (int c = 0; c < C; ++c) acc[c] += norm * e[c];
				}
			}
			#pragma unroll
			for (int c = 0; c < C; ++c) red[(g * C + c) * L + j] = acc[c];
		}
		__syncthreads();
		for (int e = threadIdx.x; e < C * L; e += blockDim.x)
		{
			const int c = e / L, jj = e - c * L;
			double v = 0.0;
			for (int gg = 0; gg < groups; ++gg) v += red[(gg * C + c) * L + jj];
			if (jj == 0)
			{
				double ac[C], bc[C];
				correlationCoefficients<CORE>(P.spin, ac, bc);
				double term = 0.0;
				#pragma unroll
				for (int cc = 0; cc < C; ++cc) if (cc == c) term = ac[cc];
				v += term * pw;
			}
			partial[((size_t)blockIdx.x * C + c) * L + jj] = outerWeight * v;
		}
	}

	__global__ void correlationSumKernel(const double *__restrict__ partial, const int *count, int entries, double *__restrict__ chi)
	{
		const int e = blockIdx.x * blockDim.x + threadIdx.x;
		if (e >= entries) return;
		double v = 0.0;
		for (int i = 0; i < count[0]; ++i) v += partial[(size_t)i * entries + e];
		chi[e] = v;
	}
}
