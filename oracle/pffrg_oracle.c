/* pffrg_oracle.c -- CPU restatement of SpinParser's pf-FRG flow-equation hot path (TEST INFRASTRUCTURE ONLY).
 *
 * See pffrg_oracle.h for the contract. Every function cites the reference lines it follows (relative to
 * /root/reference). Loop and summation order follow the reference so that an FP64 build of the reference
 * (oracle/_ref/oracle64) is reproduced to round-off of the last bit, not merely to the 1e-10 parity tolerance.
 * Compile with -ffp-contract=off and never with -ffast-math (NaN is the reference's divergence signal).
 */
#include "pffrg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------------------------------------------------------------------------------------------------------
 * Frequency mesh: src/FrequencyDiscretization.hpp:164-361. Indices are relative to the first positive mesh point;
 * the negative half is addressed with negative indices: value(-i-1) = -mesh[i] (:187-192).
 * ------------------------------------------------------------------------------------------------------------- */
double pfo_mesh_value(int nw, const double *mesh, int index)
{
	(void)nw;
	return index >= 0 ? mesh[index] : -mesh[-index - 1];
}

int pfo_mesh_greater(int nw, const double *mesh, double w);

/* FrequencyDiscretization::lesser, :253-270 (including the quirk for 0 < |w| <= mesh[0]) */
int pfo_mesh_lesser(int nw, const double *mesh, double w)
{
	if (w < 0) return -(pfo_mesh_greater(nw, mesh, -w) + 1);
	if (w <= mesh[0]) return 0;
	for (int i = 1; i < nw; ++i) if (mesh[i] > w) return i - 1;
	return nw - 1;
}

/* FrequencyDiscretization::greater, :279-296 */
int pfo_mesh_greater(int nw, const double *mesh, double w)
{
	if (w < 0) return -(pfo_mesh_lesser(nw, mesh, -w) + 1);
	if (w <= mesh[0]) return 0;
	for (int i = 1; i < nw; ++i) if (mesh[i] > w) return i;
	return nw - 1;
}

/* FrequencyDiscretization::offset, :306-316 */
int pfo_mesh_offset(int nw, const double *mesh, double w)
{
	if (w <= mesh[0]) return 0;
	for (int i = 1; i < nw; ++i) if (mesh[i] >= w) return i;
	return nw - 1;
}

/* FrequencyDiscretization::interpolateOffset, :326-351 */
void pfo_mesh_interpolate(int nw, const double *mesh, double w, int *lower, int *upper, double *bias)
{
	if (w <= mesh[0]) { *lower = 0; *upper = 0; *bias = 0.0; return; }
	for (int i = 1; i < nw; ++i)
	{
		if (mesh[i] > w)
		{
			*upper = i; *lower = i - 1;
			*bias = (w - mesh[i - 1]) / (mesh[i] - mesh[i - 1]);
			return;
		}
	}
	*lower = nw - 1; *upper = nw - 1; *bias = 0.0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Vector-valued trapezoid integrators: src/lib/Integrator.hpp:138-287 (namespace ImplicitIntegrator).
 * The integrand writes n values; evaluation order and weights as written there.
 * ------------------------------------------------------------------------------------------------------------- */
typedef void (*vec_fn)(double w, double *out, void *ctx);

static void axpy(double a, const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] += a * x[i]; }
static void vscale(double a, double *y, int n) { for (int i = 0; i < n; ++i) y[i] *= a; }
static void vadd(const double *x, double *y, int n) { for (int i = 0; i < n; ++i) y[i] += x[i]; }
#define MV(i) pfo_mesh_value(nw, mesh, (i))

/* integrateWithObscureLeftBoundary, :138-174 */
static void integrate_left(int nw, const double *mesh, double min, int max, vec_fn f, void *ctx, double *buf, double *res, int n)
{
	memset(res, 0, sizeof(double) * n);
	int umin = pfo_mesh_greater(nw, mesh, min);
	if (umin != max)
	{
		f(min, buf, ctx); axpy(MV(umin) - min, buf, res, n);
		f(MV(umin), buf, ctx); axpy(MV(umin + 1) - min, buf, res, n);
		while (++umin != max) { f(MV(umin), buf, ctx); axpy(MV(umin + 1) - MV(umin - 1), buf, res, n); }
		f(MV(umin), buf, ctx); axpy(MV(umin) - MV(umin - 1), buf, res, n);
		vscale(0.5, res, n);
	}
	else
	{
		f(min, buf, ctx); vadd(buf, res, n);
		f(MV(umin), buf, ctx); vadd(buf, res, n);
		vscale(0.5 * (MV(umin) - min), res, n);
	}
}

/* integrateWithObscureRightBoundary, :188-225 */
static void integrate_right(int nw, const double *mesh, int min, double max, vec_fn f, void *ctx, double *buf, double *res, int n)
{
	memset(res, 0, sizeof(double) * n);
	int umin = min;
	int umax = pfo_mesh_lesser(nw, mesh, max);
	if (umin != umax)
	{
		f(MV(umin), buf, ctx); axpy(MV(umin + 1) - MV(umin), buf, res, n);
		while (++umin != umax) { f(MV(umin), buf, ctx); axpy(MV(umin + 1) - MV(umin - 1), buf, res, n); }
		f(MV(umin), buf, ctx); axpy(max - MV(umin - 1), buf, res, n);
		f(max, buf, ctx); axpy(max - MV(umin), buf, res, n);
		vscale(0.5, res, n);
	}
	else
	{
		f(max, buf, ctx); vadd(buf, res, n);
		f(MV(umin), buf, ctx); vadd(buf, res, n);
		vscale(0.5 * (max - MV(umin)), res, n);
	}
}

/* integrateWithObscureBoundaries, :239-287 */
static void integrate_both(int nw, const double *mesh, double min, double max, vec_fn f, void *ctx, double *buf, double *res, int n)
{
	memset(res, 0, sizeof(double) * n);
	int umin = pfo_mesh_greater(nw, mesh, min);
	int umax = pfo_mesh_lesser(nw, mesh, max);
	if (umax >= umin)
	{
		f(min, buf, ctx); axpy(MV(umin) - min, buf, res, n);
		if (umax != umin)
		{
			f(MV(umin), buf, ctx); axpy(MV(umin + 1) - min, buf, res, n);
			while (++umin != umax) { f(MV(umin), buf, ctx); axpy(MV(umin + 1) - MV(umin - 1), buf, res, n); }
			f(MV(umin), buf, ctx); axpy(max - MV(umin - 1), buf, res, n);
		}
		else
		{
			f(MV(umin), buf, ctx); axpy(max - min, buf, res, n);
		}
		f(max, buf, ctx); axpy(max - MV(umin), buf, res, n);
		vscale(0.5, res, n);
	}
	else
	{
		f(max, buf, ctx); vadd(buf, res, n);
		f(min, buf, ctx); vadd(buf, res, n);
		vscale(0.5 * (max - min), res, n);
	}
}

/* scalar wrappers for the known-answer tests of test/test_Integrator.cpp */
typedef struct { pfo_scalar_fn f; void *ctx; } scalar_ctx;
static void scalar_adapter(double w, double *out, void *ctx) { scalar_ctx *c = (scalar_ctx *)ctx; out[0] = c->f(w, c->ctx); }
double pfo_integrate_left(int nw, const double *mesh, double min, int max_index, pfo_scalar_fn f, void *ctx)
{ scalar_ctx c = { f, ctx }; double b, r; integrate_left(nw, mesh, min, max_index, scalar_adapter, &c, &b, &r, 1); return r; }
double pfo_integrate_right(int nw, const double *mesh, int min_index, double max, pfo_scalar_fn f, void *ctx)
{ scalar_ctx c = { f, ctx }; double b, r; integrate_right(nw, mesh, min_index, max, scalar_adapter, &c, &b, &r, 1); return r; }
double pfo_integrate_both(int nw, const double *mesh, double min, double max, pfo_scalar_fn f, void *ctx)
{ scalar_ctx c = { f, ctx }; double b, r; integrate_both(nw, mesh, min, max, scalar_adapter, &c, &b, &r, 1); return r; }

/* node counts of the three integrators */
int pfo_node_count(const pfo_problem *p, double cutoff, double x)
{
	int nw = p->nw; const double *mesh = p->mesh;
	int n = 1;
	if (x > 2.0 * cutoff) n += 1;
	if (-(x + cutoff) > -mesh[nw - 1])
	{
		int umax = pfo_mesh_lesser(nw, mesh, -(x + cutoff));
		n += (umax != -nw) ? (umax + nw) + 2 : 2;
	}
	if (x - cutoff > cutoff)
	{
		int umin = pfo_mesh_greater(nw, mesh, cutoff - x), umax = pfo_mesh_lesser(nw, mesh, -cutoff);
		n += (umax >= umin) ? (umax - umin) + 3 : 2;
	}
	if (cutoff < mesh[nw - 1])
	{
		int umin = pfo_mesh_greater(nw, mesh, cutoff);
		n += (umin != nw - 1) ? (nw - 1 - umin) + 2 : 2;
	}
	return n;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Self energy: {SU2,XYZ,TRI}VertexSingleParticle::getValue, src/SU2/SU2VertexSingleParticle.hpp:73-87 (odd, clamped lerp)
 * ------------------------------------------------------------------------------------------------------------- */
static double v2_value(const pfo_problem *p, const double *v2, double w)
{
	int lo, up; double bias, sign = 1.0;
	if (w < 0) { w = -w; sign = -1.0; }
	pfo_mesh_interpolate(p->nw, p->mesh, w, &lo, &up, &bias);
	return sign * ((1 - bias) * v2[lo] + bias * v2[up]);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Access buffers.
 * SU2/XYZ: src/SU2/SU2VertexTwoParticle.hpp:399-490 (4 supports), :500-557 (8 supports), :616-632
 * TRI:     src/TRI/TRIVertexTwoParticle.hpp:401-504, :643-666
 * `off` is the row index su*nw+t (the reference's frequencyOffsets divided by the row length).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct
{
	int n;              /* 4 or 8 supports */
	int off[8];
	double w[8];
	int swapped[8];     /* s index < u index: stored entry is the s<->u mirrored one */
	int exchange;       /* siteExchange / pairExchange */
	int tz;             /* TRI: number of zeta_mu*zeta_nu factors (mod 2) */
} access_buffer;

enum { CH_S = 0, CH_T = 1, CH_U = 2, CH_NONE = 3 };

static void ab_offset(const pfo_problem *p, access_buffer *ab, int k, int so, int to, int uo)
{
	if (so < uo) { ab->swapped[k] = 1; ab->off[k] = (uo * (uo + 1) / 2 + so) * p->nw + to; }
	else { ab->swapped[k] = 0; ab->off[k] = (so * (so + 1) / 2 + uo) * p->nw + to; }
}

static access_buffer make_access_buffer(const pfo_problem *p, double s, double t, double u, int channel)
{
	access_buffer ab; memset(&ab, 0, sizeof(ab));
	int nw = p->nw; const double *mesh = p->mesh;
	if (p->core == PFO_TRI)
	{
		/* TRIVertexTwoParticle.hpp:408-443 */
		if (s < 0) { s = -s; ab.exchange = !ab.exchange; }
		if (t < 0) { t = -t; ab.tz ^= 1; }
		if (u < 0) { u = -u; ab.exchange = !ab.exchange; ab.tz ^= 1; }
	}
	else
	{
		/* SU2VertexTwoParticle.hpp:406-427 */
		if (s < 0 && u < 0) { s = -s; u = -u; }
		else
		{
			if (s < 0) { s = -s; ab.exchange = 1; }
			else if (u < 0) { u = -u; ab.exchange = 1; }
		}
		if (t < 0) t = -t;
	}
	int ls, us, lt, ut, lu, uu; double bs, bt, bu;
	if (channel == CH_S)
	{
		int es = pfo_mesh_offset(nw, mesh, s);
		pfo_mesh_interpolate(nw, mesh, t, &lt, &ut, &bt);
		pfo_mesh_interpolate(nw, mesh, u, &lu, &uu, &bu);
		ab.n = 4;
		ab.w[0] = (1 - bu) * (1 - bt); ab_offset(p, &ab, 0, es, lt, lu);
		ab.w[1] = (1 - bu) * bt;       ab_offset(p, &ab, 1, es, ut, lu);
		ab.w[2] = bu * (1 - bt);       ab_offset(p, &ab, 2, es, lt, uu);
		ab.w[3] = bu * bt;             ab_offset(p, &ab, 3, es, ut, uu);
	}
	else if (channel == CH_T)
	{
		int et = pfo_mesh_offset(nw, mesh, t);
		pfo_mesh_interpolate(nw, mesh, s, &ls, &us, &bs);
		pfo_mesh_interpolate(nw, mesh, u, &lu, &uu, &bu);
		ab.n = 4;
		ab.w[0] = (1 - bu) * (1 - bs); ab_offset(p, &ab, 0, ls, et, lu);
		ab.w[1] = (1 - bu) * bs;       ab_offset(p, &ab, 1, us, et, lu);
		ab.w[2] = bu * (1 - bs);       ab_offset(p, &ab, 2, ls, et, uu);
		ab.w[3] = bu * bs;             ab_offset(p, &ab, 3, us, et, uu);
	}
	else if (channel == CH_U)
	{
		int eu = pfo_mesh_offset(nw, mesh, u);
		pfo_mesh_interpolate(nw, mesh, s, &ls, &us, &bs);
		pfo_mesh_interpolate(nw, mesh, t, &lt, &ut, &bt);
		ab.n = 4;
		ab.w[0] = (1 - bt) * (1 - bs); ab_offset(p, &ab, 0, ls, lt, eu);
		ab.w[1] = (1 - bt) * bs;       ab_offset(p, &ab, 1, us, lt, eu);
		ab.w[2] = bt * (1 - bs);       ab_offset(p, &ab, 2, ls, ut, eu);
		ab.w[3] = bt * bs;             ab_offset(p, &ab, 3, us, ut, eu);
	}
	else
	{
		pfo_mesh_interpolate(nw, mesh, s, &ls, &us, &bs);
		pfo_mesh_interpolate(nw, mesh, t, &lt, &ut, &bt);
		pfo_mesh_interpolate(nw, mesh, u, &lu, &uu, &bu);
		ab.n = 8;
		ab.w[0] = (1 - bt) * (1 - bs) * (1 - bu); ab_offset(p, &ab, 0, ls, lt, lu);
		ab.w[1] = (1 - bt) * bs * (1 - bu);       ab_offset(p, &ab, 1, us, lt, lu);
		ab.w[2] = bt * (1 - bs) * (1 - bu);       ab_offset(p, &ab, 2, ls, ut, lu);
		ab.w[3] = bt * bs * (1 - bu);             ab_offset(p, &ab, 3, us, ut, lu);
		ab.w[4] = (1 - bt) * (1 - bs) * bu;       ab_offset(p, &ab, 4, ls, lt, uu);
		ab.w[5] = (1 - bt) * bs * bu;             ab_offset(p, &ab, 5, us, lt, uu);
		ab.w[6] = bt * (1 - bs) * bu;             ab_offset(p, &ab, 6, ls, ut, uu);
		ab.w[7] = bt * bs * bu;                   ab_offset(p, &ab, 7, us, ut, uu);
	}
	return ab;
}

static double zeta(int c) { return c <= 2 ? -1.0 : 1.0; } /* TRIVertexTwoParticle.hpp:674-677 */

/* sign of support k for output channel c (SU2: 0 spin, 1 density; XYZ: 0..2 spin, 3 density; TRI: 4*mu+nu) */
static double ab_sign(const pfo_problem *p, const access_buffer *ab, int k, int c)
{
	if (p->core == PFO_SU2) return (c == 1 && ab->swapped[k]) ? -1.0 : 1.0;
	if (p->core == PFO_XYZ) return (c == 3 && ab->swapped[k]) ? -1.0 : 1.0;
	int mu = c / 4, nu = c % 4;
	double sgn = 1.0;
	if (ab->tz) sgn *= zeta(mu) * zeta(nu);
	if (ab->swapped[k]) sgn *= ab->exchange ? -zeta(mu) : -zeta(nu);
	return sgn;
}

/* element address of (row, stored channel, rid): array pointer and index */
static double v4_elem(const pfo_problem *p, const double *const *v4, int row, int c, int rid)
{
	if (p->core == PFO_TRI) return v4[0][((long)row * 16 + c) * p->L + rid];
	return v4[c][(long)row * p->L + rid];
}

/* stored channel that output channel c of site descriptor (perm) reads */
static int stored_channel(const pfo_problem *p, const access_buffer *ab, int c, const int *perm)
{
	if (p->core == PFO_SU2) return c;
	if (p->core == PFO_XYZ) return c < 3 ? perm[c] : 3; /* XYZVertexTwoParticle.hpp:401-404 */
	int mu = c / 4, nu = c % 4;                         /* TRIVertexTwoParticle.hpp:378-382 */
	int m = ab->exchange ? nu : mu, n = ab->exchange ? mu : nu;
	if (m < 3) m = perm[m];
	if (n < 3) n = perm[n];
	return 4 * m + n;
}

/* getValueSuperbundle: SU2VertexTwoParticle.hpp:369-387, XYZVertexTwoParticle.hpp:385-407, TRIVertexTwoParticle.hpp:364-390.
 * out[c*L + j] */
static void gather(const pfo_problem *p, const double *const *v4, const access_buffer *ab, double *out)
{
	int L = p->L, C = pfo_num_channels(p->core);
	const int *rid = ab->exchange ? p->inv_rid : p->sites_rid;
	const int *perm = ab->exchange ? p->inv_perm : p->sites_perm;
	memset(out, 0, sizeof(double) * C * L);
	for (int k = 0; k < ab->n; ++k)
		for (int c = 0; c < C; ++c)
		{
			double sw = ab_sign(p, ab, k, c) * ab->w[k];
			for (int j = 0; j < L; ++j) out[c * L + j] += sw * v4_elem(p, v4, ab->off[k], stored_channel(p, ab, c, perm + 3 * j), rid[j]);
		}
}

/* getValueLocal: SU2VertexTwoParticle.hpp:347-360, XYZ :350-373, TRI :348-357 (site 0, no spin permutation) */
static double gather_local(const pfo_problem *p, const double *const *v4, const access_buffer *ab, int c)
{
	static const int ident[3] = { 0, 1, 2 };
	double value = 0.0;
	int sc = stored_channel(p, ab, c, ident);
	/* TRI: getValueLocal swaps (s1, s2) under pairExchange BEFORE it indexes the sign table (TRIVertexTwoParticle.hpp:349-353),
	 * unlike getValueSuperbundle (:385), which indexes it with the unswapped output pair */
	int csign = (p->core == PFO_TRI && ab->exchange) ? 4 * (c % 4) + c / 4 : c;
	for (int k = 0; k < ab->n; ++k) value += ab_sign(p, ab, k, csign) * ab->w[k] * v4_elem(p, v4, ab->off[k], sc, 0);
	return value;
}

/* getValue(i1=0, i2=j, s,t,u, channel None): 8-support trilinear access with explicit nesting,
 * SU2VertexTwoParticle.hpp:185-215,270-297; XYZ :188-213,273-300; TRI :203-226,...
 * Only diagonal channels are needed by the self-energy flow, for which spin permutations act trivially on the
 * density channel and the local terms use the identity pair (0,0). */
static double v4_value_none(const pfo_problem *p, const double *const *v4, int fwd_rid, int inv_rid, double s, double t, double u, int c)
{
	access_buffer ab = make_access_buffer(p, s, t, u, CH_NONE);
	int site = ab.exchange ? inv_rid : fwd_rid;
	double v[8];
	for (int k = 0; k < 8; ++k) v[k] = ab_sign(p, &ab, k, c) * v4_elem(p, v4, ab.off[k], c, site);
	/* recover the biases from the weights is not possible; recompute them as the reference does */
	int l, h; double bs, bt, bu;
	double as = fabs(s), at = fabs(t), au = fabs(u);
	pfo_mesh_interpolate(p->nw, p->mesh, as, &l, &h, &bs);
	pfo_mesh_interpolate(p->nw, p->mesh, at, &l, &h, &bt);
	pfo_mesh_interpolate(p->nw, p->mesh, au, &l, &h, &bu);
	/* support order of make_access_buffer: k = (u?4:0) + (t?2:0) + (s?1:0) */
	return (1 - bu) * ((1 - bt) * ((1 - bs) * v[0] + bs * v[1]) + bt * ((1 - bs) * v[2] + bs * v[3]))
	     + bu * ((1 - bt) * ((1 - bs) * v[4] + bs * v[5]) + bt * ((1 - bs) * v[6] + bs * v[7]));
}

int pfo_num_arrays(int core) { return core == PFO_SU2 ? 2 : core == PFO_XYZ ? 4 : 1; }
int pfo_num_channels(int core) { return core == PFO_SU2 ? 2 : core == PFO_XYZ ? 4 : 16; }

/* ---------------------------------------------------------------------------------------------------------------
 * Self-energy flow: SU2FrgCore.cpp:139-169, XYZFrgCore.cpp:164-193, TRIFrgCore.cpp:122-151
 * ------------------------------------------------------------------------------------------------------------- */
void pfo_v2_flow(const pfo_problem *p, double cutoff, const double *v2, const double *const *v4, double *v2flow)
{
	int dens = p->core == PFO_SU2 ? 1 : p->core == PFO_XYZ ? 3 : 15;
	#pragma omp parallel for schedule(static)
	for (int it = 0; it < p->nw; ++it)
	{
		double w = p->mesh[it], value = 0.0, sum = 0.0;
		for (int j = 0; j < p->nrange; ++j)
		{
			sum += v4_value_none(p, v4, p->rng_fwd_rid[j], p->rng_inv_rid[j], w + cutoff, 0.0, w - cutoff, dens);
			sum -= v4_value_none(p, v4, p->rng_fwd_rid[j], p->rng_inv_rid[j], w - cutoff, 0.0, w + cutoff, dens);
		}
		if (p->core == PFO_SU2)
		{
			value -= 4.0 * p->spin_length * sum;
			value += 0.75 * (v4_value_none(p, v4, 0, 0, w + cutoff, w - cutoff, 0.0, 0) - v4_value_none(p, v4, 0, 0, w - cutoff, w + cutoff, 0.0, 0));
			value += (v4_value_none(p, v4, 0, 0, w + cutoff, w - cutoff, 0.0, 1) - v4_value_none(p, v4, 0, 0, w - cutoff, w + cutoff, 0.0, 1));
		}
		else
		{
			value -= 2.0 * sum;
			for (int k = 0; k < 4; ++k)
			{
				int c = p->core == PFO_XYZ ? k : 5 * k;
				value += v4_value_none(p, v4, 0, 0, w + cutoff, w - cutoff, 0.0, c) - v4_value_none(p, v4, 0, 0, w - cutoff, w + cutoff, 0.0, c);
			}
		}
		value /= (2.0 * (double)M_PI * (cutoff + v2_value(p, v2, cutoff)));
		v2flow[it] = value;
	}
}

/* ---------------------------------------------------------------------------------------------------------------
 * Two-particle vertex flow
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct
{
	const pfo_problem *p;
	const double *const *v4;
	const double *v2, *v2flow;
	double s, t, u, w1p, w1, w2p, w2;
	double *sb[4];      /* stackBuffers */
	double *rpa;        /* bufferRPA */
	int channel;        /* which kernel: CH_S, CH_T, CH_U */
	double katanin_sign;
	double x;           /* transfer frequency of the channel */
} item_ctx;

/* y[c_out] += a * x1[c1] * x2[c2] over all sites (ValueBundle::multAdd(a, x, z): a*x[i]*z[i], src/lib/ValueBundle.hpp:101-105) */
static void mad3(double *y, double a, const double *x1, const double *x2, int L) { for (int i = 0; i < L; ++i) y[i] += a * x1[i] * x2[i]; }
static void msub3(double *y, double a, const double *x1, const double *x2, int L) { for (int i = 0; i < L; ++i) y[i] -= a * x1[i] * x2[i]; }
static void mad2(double *y, const double *x1, const double *x2, int L) { for (int i = 0; i < L; ++i) y[i] += x1[i] * x2[i]; }
static void msub2(double *y, const double *x1, const double *x2, int L) { for (int i = 0; i < L; ++i) y[i] -= x1[i] * x2[i]; }
static void mad1(double *y, double a, const double *x, int L) { for (int i = 0; i < L; ++i) y[i] += a * x[i]; }
static void msub1(double *y, double a, const double *x, int L) { for (int i = 0; i < L; ++i) y[i] -= a * x[i]; }

#ifdef PFO_HAVE_TRI_TABLES
/* Term tables of the TRI core, generated at build time from src/TRI/TRIFrgCore.cpp:198-2902 by oracle/gen_tri_terms.py
 * into oracle/_ref/ (never committed): {out, sign, first, second} per region, RPA additionally which index is permuted. */
#include "_ref/tri_terms.inc"
#endif

static void gather4(item_ctx *c, const access_buffer *ab)
{
	for (int k = 0; k < 4; ++k) gather(c->p, c->v4, &ab[k], c->sb[k]);
}

/* integralKernelS: SU2FrgCore.cpp:199-227, XYZFrgCore.cpp:223-275, TRIFrgCore.cpp:181-712 */
static void kernel_s(item_ctx *c, double wp, double *ret)
{
	const pfo_problem *p = c->p; int L = p->L, C = pfo_num_channels(p->core);
	double s = c->s, w1 = c->w1, w2 = c->w2, w1p = c->w1p, w2p = c->w2p;
	access_buffer ab[4];
	if (p->core == PFO_TRI)
	{
		ab[0] = make_access_buffer(p, s, w2 + wp, w1 + wp, CH_S);
		ab[1] = make_access_buffer(p, s, -w2p - wp, w1p + wp, CH_S);
		ab[2] = make_access_buffer(p, s, -w1 - wp, -w2 - wp, CH_S);
		ab[3] = make_access_buffer(p, s, w1p + wp, -w2p - wp, CH_S);
	}
	else
	{
		ab[0] = make_access_buffer(p, s, -w1 - wp, -w2 - wp, CH_S);
		ab[1] = make_access_buffer(p, s, w1p + wp, -w2p - wp, CH_S);
		ab[2] = make_access_buffer(p, s, w2 + wp, w1 + wp, CH_S);
		ab[3] = make_access_buffer(p, s, -w2p - wp, w1p + wp, CH_S);
	}
	gather4(c, ab);
	memset(ret, 0, sizeof(double) * C * L);
	double **sb = c->sb;
	if (p->core == PFO_SU2)
	{
		double *rs = ret, *rd = ret + L;
		#define S_(k) (sb[k])
		#define D_(k) (sb[k] + L)
		msub3(rs, 0.5, S_(0), S_(1), L); msub3(rs, 0.5, S_(2), S_(3), L);
		mad2(rs, D_(0), S_(1), L); mad2(rs, D_(2), S_(3), L);
		mad2(rs, S_(0), D_(1), L); mad2(rs, S_(2), D_(3), L);
		mad3(rd, 3.0 / 16.0, S_(0), S_(1), L); mad3(rd, 3.0 / 16.0, S_(2), S_(3), L);
		mad2(rd, D_(0), D_(1), L); mad2(rd, D_(2), D_(3), L);
	}
	else if (p->core == PFO_XYZ)
	{
		#define B_(k, ch) (sb[k] + (ch) * L)
		static const int other[3][2] = { { 2, 1 }, { 2, 0 }, { 0, 1 } }; /* (Z,Y), (Z,X), (X,Y) : XYZFrgCore.cpp:240-265 */
		for (int a = 0; a < 3; ++a)
			for (int pr = 0; pr < 4; pr += 2)
			{
				mad2(ret + a * L, B_(pr, 3), B_(pr + 1, a), L);
				mad2(ret + a * L, B_(pr, a), B_(pr + 1, 3), L);
				msub2(ret + a * L, B_(pr, other[a][0]), B_(pr + 1, other[a][1]), L);
				msub2(ret + a * L, B_(pr, other[a][1]), B_(pr + 1, other[a][0]), L);
			}
		for (int pr = 0; pr < 4; pr += 2)
		{
			mad2(ret + 3 * L, B_(pr, 3), B_(pr + 1, 3), L);
			mad2(ret + 3 * L, B_(pr, 0), B_(pr + 1, 0), L);
			mad2(ret + 3 * L, B_(pr, 1), B_(pr + 1, 1), L);
			mad2(ret + 3 * L, B_(pr, 2), B_(pr + 1, 2), L);
		}
	}
#ifdef PFO_HAVE_TRI_TABLES
	else
	{
		for (int i = 0; i < TRI_N_PPLADDER; ++i)
		{
			const tri_term *t = &tri_ppladder[i];
			if (t->sign > 0) mad2(ret + t->out * L, B_(t->buf1, t->c1), B_(t->buf2, t->c2), L);
			else msub2(ret + t->out * L, B_(t->buf1, t->c1), B_(t->buf2, t->c2), L);
		}
	}
#endif
}

/* integralKernelU: SU2FrgCore.cpp:305-334, XYZFrgCore.cpp:419-472, TRIFrgCore.cpp:2371-2903 */
static void kernel_u(item_ctx *c, double wp, double *ret)
{
	const pfo_problem *p = c->p; int L = p->L, C = pfo_num_channels(p->core);
	double u = c->u, w1 = c->w1, w2 = c->w2, w1p = c->w1p, w2p = c->w2p;
	access_buffer ab[4];
	ab[0] = make_access_buffer(p, w1 + wp, wp - w2p, u, CH_U);
	ab[1] = make_access_buffer(p, w1p + wp, w2 - wp, u, CH_U);
	ab[2] = make_access_buffer(p, w2p - wp, -w1 - wp, u, CH_U);
	ab[3] = make_access_buffer(p, w2 - wp, w1p + wp, u, CH_U);
	gather4(c, ab);
	memset(ret, 0, sizeof(double) * C * L);
	double **sb = c->sb;
	if (p->core == PFO_SU2)
	{
		double *rs = ret, *rd = ret + L;
		mad3(rs, 0.5, S_(0), S_(1), L); mad3(rs, 0.5, S_(2), S_(3), L);
		mad2(rs, S_(0), D_(1), L); mad2(rs, S_(2), D_(3), L);
		mad2(rs, D_(0), S_(1), L); mad2(rs, D_(2), S_(3), L);
		mad3(rd, 3.0 / 16.0, S_(0), S_(1), L); mad3(rd, 3.0 / 16.0, S_(2), S_(3), L);
		mad2(rd, D_(0), D_(1), L); mad2(rd, D_(2), D_(3), L);
	}
	else if (p->core == PFO_XYZ)
	{
		static const int other[3][2] = { { 2, 1 }, { 2, 0 }, { 0, 1 } }; /* XYZFrgCore.cpp:437-462 */
		for (int a = 0; a < 3; ++a)
			for (int pr = 0; pr < 4; pr += 2)
			{
				msub2(ret + a * L, B_(pr, 3), B_(pr + 1, a), L);
				msub2(ret + a * L, B_(pr, a), B_(pr + 1, 3), L);
				msub2(ret + a * L, B_(pr, other[a][0]), B_(pr + 1, other[a][1]), L);
				msub2(ret + a * L, B_(pr, other[a][1]), B_(pr + 1, other[a][0]), L);
			}
		for (int pr = 0; pr < 4; pr += 2)
		{
			msub2(ret + 3 * L, B_(pr, 3), B_(pr + 1, 3), L);
			msub2(ret + 3 * L, B_(pr, 0), B_(pr + 1, 0), L);
			msub2(ret + 3 * L, B_(pr, 1), B_(pr + 1, 1), L);
			msub2(ret + 3 * L, B_(pr, 2), B_(pr + 1, 2), L);
		}
	}
#ifdef PFO_HAVE_TRI_TABLES
	else
	{
		for (int i = 0; i < TRI_N_PHLADDER; ++i)
		{
			const tri_term *t = &tri_phladder[i];
			if (t->sign > 0) mad2(ret + t->out * L, B_(t->buf1, t->c1), B_(t->buf2, t->c2), L);
			else msub2(ret + t->out * L, B_(t->buf1, t->c1), B_(t->buf2, t->c2), L);
		}
	}
#endif
}

/* integralKernelT: SU2FrgCore.cpp:229-303, XYZFrgCore.cpp:277-417, TRIFrgCore.cpp:714-2369 */
static void kernel_t(item_ctx *c, double wp, double *ret)
{
	const pfo_problem *p = c->p; int L = p->L, C = pfo_num_channels(p->core);
	double t = c->t, w1 = c->w1, w2 = c->w2, w1p = c->w1p, w2p = c->w2p;
	access_buffer ab[4];
	ab[0] = make_access_buffer(p, w1 - wp, t, w1p + wp, CH_T);
	ab[1] = make_access_buffer(p, w2p - wp, t, -w2 - wp, CH_T);
	ab[2] = make_access_buffer(p, w1p + wp, t, w1 - wp, CH_T);
	ab[3] = make_access_buffer(p, w2 + wp, t, wp - w2p, CH_T);
	gather4(c, ab);
	memset(ret, 0, sizeof(double) * C * L);
	double **sb = c->sb; double *rpa = c->rpa;
	memset(rpa, 0, sizeof(double) * C * L);

	/* local (site 0) buffers; SU2/XYZ naming: 4 = chalice B, 5 = inverse chalice A, 6 = chalice B', 7 = inverse chalice A' */
	access_buffer cb = make_access_buffer(p, w2p - wp, -w2 - wp, t, CH_U);
	access_buffer ica = make_access_buffer(p, w1 - wp, -w1p - wp, -t, CH_U);
	access_buffer cb2 = make_access_buffer(p, w2 + wp, wp - w2p, t, CH_U);
	access_buffer ica2 = make_access_buffer(p, w1p + wp, wp - w1, -t, CH_U);

	if (p->core == PFO_SU2)
	{
		for (int rid = 0; rid < L; ++rid)
		{
			for (int i = p->ov_off[rid]; i < p->ov_off[rid + 1]; ++i) rpa[rid] += S_(2)[p->ov_rid1[i]] * S_(3)[p->ov_rid2[i]];
			for (int i = p->ov_off[rid]; i < p->ov_off[rid + 1]; ++i) rpa[L + rid] += D_(2)[p->ov_rid1[i]] * D_(3)[p->ov_rid2[i]];
		}
		mad1(ret, 2.0 * p->spin_length, rpa, L);
		mad1(ret + L, 8.0 * p->spin_length, rpa + L, L);

		const access_buffer *loc[4] = { &cb, &ica, &cb2, &ica2 };
		double vs[4], vd[4];
		for (int k = 0; k < 4; ++k) { vs[k] = gather_local(p, c->v4, loc[k], 0); vd[k] = gather_local(p, c->v4, loc[k], 1); }
		for (int k = 0; k < 4; ++k) { msub1(ret, vd[k], S_(k), L); mad1(ret, 0.25 * vs[k], S_(k), L); }
		for (int k = 0; k < 4; ++k) { msub1(ret + L, vd[k], D_(k), L); msub1(ret + L, 0.75 * vs[k], D_(k), L); }
	}
	else if (p->core == PFO_XYZ)
	{
		for (int rid = 0; rid < L; ++rid)
		{
			for (int a = 0; a < 3; ++a)
				for (int i = p->ov_off[rid]; i < p->ov_off[rid + 1]; ++i)
					rpa[a * L + rid] += B_(0, p->ov_perm1[3 * i + a])[p->ov_rid1[i]] * B_(1, p->ov_perm2[3 * i + a])[p->ov_rid2[i]];
			for (int i = p->ov_off[rid]; i < p->ov_off[rid + 1]; ++i) rpa[3 * L + rid] += B_(0, 3)[p->ov_rid1[i]] * B_(1, 3)[p->ov_rid2[i]];
		}
		mad1(ret, 4.0, rpa, 4 * L);

		const access_buffer *loc[4] = { &cb, &ica, &cb2, &ica2 };
		double v[4][4];
		for (int k = 0; k < 4; ++k) for (int q = 0; q < 4; ++q) v[k][q] = gather_local(p, c->v4, loc[k], q);
		for (int a = 0; a < 3; ++a)
			for (int k = 0; k < 4; ++k)
			{
				int o1 = (a + 1) % 3, o2 = (a + 2) % 3; /* the two other spin components enter with + (XYZFrgCore.cpp:350-400) */
				msub1(ret + a * L, v[k][3], B_(k, a), L);
				mad1(ret + a * L, v[k][o1], B_(k, a), L);
				mad1(ret + a * L, v[k][o2], B_(k, a), L);
				msub1(ret + a * L, v[k][a], B_(k, a), L);
			}
		for (int k = 0; k < 4; ++k)
		{
			msub1(ret + 3 * L, v[k][3], B_(k, 3), L);
			msub1(ret + 3 * L, v[k][0], B_(k, 3), L);
			msub1(ret + 3 * L, v[k][1], B_(k, 3), L);
			msub1(ret + 3 * L, v[k][2], B_(k, 3), L);
		}
	}
#ifdef PFO_HAVE_TRI_TABLES
	else
	{
		for (int rid = 0; rid < L; ++rid)
			for (int q = 0; q < TRI_N_RPA; ++q)
			{
				const tri_rpa_term *t = &tri_rpa[q];
				for (int i = p->ov_off[rid]; i < p->ov_off[rid + 1]; ++i)
				{
					int c1 = t->c1, c2 = t->c2;
					if (t->p1a >= 0) c1 = 4 * p->ov_perm1[3 * i + t->p1a] + (t->p1b >= 0 ? p->ov_perm1[3 * i + t->p1b] : (t->p1b == -2 ? 3 : 0));
					else if (t->p1b >= 0) c1 = 12 + p->ov_perm1[3 * i + t->p1b];
					if (t->p2a >= 0) c2 = 4 * p->ov_perm2[3 * i + t->p2a] + (t->p2b >= 0 ? p->ov_perm2[3 * i + t->p2b] : (t->p2b == -2 ? 3 : 0));
					else if (t->p2b >= 0) c2 = 12 + p->ov_perm2[3 * i + t->p2b];
					rpa[t->out * L + rid] += t->sign * 2 * B_(t->buf1, c1)[p->ov_rid1[i]] * B_(t->buf2, c2)[p->ov_rid2[i]];
				}
			}
		vadd(rpa, ret, 16 * L);
		/* TRI naming: ab4 = chalice B (with buffer 0), ab5 = chalice B' (with buffer 2), ab6/ab7 = inverse chalice A (with buffers 1/3) */
		double loc4[16], loc5[16], loc6[16], loc7[16];
		for (int q = 0; q < 16; ++q) { loc4[q] = gather_local(p, c->v4, &cb, q); loc5[q] = gather_local(p, c->v4, &cb2, q); loc6[q] = gather_local(p, c->v4, &ica, q); loc7[q] = gather_local(p, c->v4, &ica2, q); }
		const double *locs[4] = { loc4, loc5, loc6, loc7 };
		for (int i = 0; i < TRI_N_CHALICE; ++i)
		{
			const tri_term *t = &tri_chalice[i];
			if (t->sign > 0) mad1(ret + t->out * L, locs[t->buf2 - 4][t->c2], B_(t->buf1, t->c1), L);
			else msub1(ret + t->out * L, locs[t->buf2 - 4][t->c2], B_(t->buf1, t->c1), L);
		}
		for (int i = 0; i < TRI_N_INVCHALICE; ++i)
		{
			const tri_term *t = &tri_invchalice[i];
			if (t->sign > 0) mad1(ret + t->out * L, locs[t->buf2 - 4][t->c2], B_(t->buf1, t->c1), L);
			else msub1(ret + t->out * L, locs[t->buf2 - 4][t->c2], B_(t->buf1, t->c1), L);
		}
	}
#endif
}

static double bubble(const item_ctx *c, double w1, double w2) /* SU2FrgCore.cpp:337-340 */
{
	return 1.0 / ((w1 + v2_value(c->p, c->v2, w1)) * (w2 + v2_value(c->p, c->v2, w2)));
}

static double katanin(const item_ctx *c, double w1, double w2) /* SU2FrgCore.cpp:343-347 */
{
	double d = w1 + v2_value(c->p, c->v2, w1);
	return v2_value(c->p, c->v2flow, w1) / (d * d * (w2 + v2_value(c->p, c->v2, w2)));
}

static void run_kernel(item_ctx *c, double wp, double *ret)
{
	if (c->channel == CH_S) kernel_s(c, wp, ret);
	else if (c->channel == CH_T) kernel_t(c, wp, ret);
	else kernel_u(c, wp, ret);
}

/* integralKernel{S,T,U}Katanin: SU2FrgCore.cpp:374-376 */
static void katanin_integrand(double wp, double *out, void *ctx)
{
	item_ctx *c = (item_ctx *)ctx;
	run_kernel(c, wp, out);
	vscale(c->katanin_sign * katanin(c, wp, c->x + wp), out, pfo_num_channels(c->p->core) * c->p->L);
}

void pfo_v4_flow(const pfo_problem *p, double cutoff, const double *v2, const double *v2flow, const double *const *v4,
                 const int *items, int n_items, double *const *flow)
{
	int nw = p->nw, L = p->L, C = pfo_num_channels(p->core), n = C * L;
	const double *mesh = p->mesh;
	/* SU2: the u channel enters with a minus sign (SU2FrgCore.cpp:366,370,376); XYZ/TRI kernels carry it themselves */
	double usign = p->core == PFO_SU2 ? -1.0 : 1.0;

	#pragma omp parallel
	{
		double *mem = (double *)malloc(sizeof(double) * n * 8);
		item_ctx c; memset(&c, 0, sizeof(c));
		c.p = p; c.v4 = v4; c.v2 = v2; c.v2flow = v2flow;
		double *buffer1 = mem, *buffer2 = mem + n, *value = mem + 2 * n;
		c.rpa = mem + 3 * n;
		for (int k = 0; k < 4; ++k) c.sb[k] = mem + (4 + k) * n;

		#pragma omp for schedule(guided)
		for (int ii = 0; ii < n_items; ++ii)
		{
			int it = items ? items[ii] : ii;
			/* expandIterator, SU2VertexTwoParticle.hpp:136-158 */
			int su = it / nw, so = 0;
			while ((so + 1) * (so + 2) / 2 <= su) ++so;
			int uo = su - so * (so + 1) / 2;
			double s = mesh[so], t = mesh[it % nw], u = mesh[uo];
			c.s = s; c.t = t; c.u = u;
			c.w1p = 0.5 * (s + t + u); c.w1 = 0.5 * (s - t + u); c.w2p = 0.5 * (s - t - u); c.w2 = 0.5 * (s + t - u);
			memset(value, 0, sizeof(double) * n);

			/* conventional contribution, SU2FrgCore.cpp:351-371 */
			const double xs[3] = { s, t, u };
			for (int ch = 0; ch < 3; ++ch)
			{
				double x = xs[ch], sg = ch == CH_U ? usign : 1.0;
				c.channel = ch;
				run_kernel(&c, cutoff, buffer1);
				axpy(sg * bubble(&c, cutoff, cutoff + x), buffer1, value, n);
				if (x > 2.0 * cutoff)
				{
					run_kernel(&c, -cutoff, buffer1);
					axpy(sg * bubble(&c, cutoff, cutoff - x), buffer1, value, n);
				}
			}
			/* Katanin contribution, SU2FrgCore.cpp:378-424 */
			for (int ch = 0; ch < 3; ++ch)
			{
				double x = xs[ch];
				c.channel = ch; c.x = x; c.katanin_sign = ch == CH_U ? usign : 1.0;
				if (-(x + cutoff) > -mesh[nw - 1])
				{
					integrate_right(nw, mesh, -nw, -(x + cutoff), katanin_integrand, &c, buffer1, buffer2, n);
					vadd(buffer2, value, n);
				}
				if (x - cutoff > cutoff)
				{
					integrate_both(nw, mesh, cutoff - x, -cutoff, katanin_integrand, &c, buffer1, buffer2, n);
					vadd(buffer2, value, n);
				}
				if (cutoff < mesh[nw - 1])
				{
					integrate_left(nw, mesh, cutoff, nw - 1, katanin_integrand, &c, buffer1, buffer2, n);
					vadd(buffer2, value, n);
				}
			}
			/* prefactor and scatter, SU2FrgCore.cpp:427-430 */
			for (int i = 0; i < n; ++i) value[i] /= 2.0 * (double)M_PI;
			if (p->core == PFO_TRI) memcpy(flow[0] + (long)it * n, value, sizeof(double) * n);
			else for (int ch = 0; ch < C; ++ch) memcpy(flow[ch] + (long)it * L, value + ch * L, sizeof(double) * L);
		}
		free(mem);
	}
}

void pfo_euler(double *x, const double *flow, long n, double cutoff, double new_cutoff)
{
	double step = new_cutoff - cutoff;
	#pragma omp parallel for schedule(static)
	for (long i = 0; i < n; ++i) x[i] += step * flow[i];
}
