/*
 * cvtx_oracle_remesh.c -- CPU ORACLE for particle redistribution onto a grid and
 * Pedrizzetti relaxation.  TEST INFRASTRUCTURE ONLY (see cvtx_oracle.h): loaded by tests/,
 * smoke() and the cpu_baseline leg of tools/remesh_bench.py, never by the product.
 *
 * A sequential restatement of the reference algorithm in the reference's arithmetic: FP32
 * shares, each node's shares added in FP32 in the order the particles arrive (what the
 * reference's tree does with one thread, src/GridParticleOcttree.cpp:76-135), FP32 sums
 * for the mean strength and the vorticity deficit, the same 1024-bin threshold search.
 * The tree itself is not restated: a node set keyed by grid index, written out in the
 * order the reference's depth-first flatten produces (ascending Morton code, x lowest),
 * is the same function of the input.
 *
 * PARITY IS PINNED against the reference's own implementation compiled from
 * /root/reference (oracle/_ref; its two tree units go through msvc_shim/ref_tree_msvc.h
 * because the reference's g++ branch is an `assert(false); / * TO DO * /` stub):
 * tests/test_remesh.py compares node sets, order, strengths and counts, live and through
 * tests/golden/reference_remesh.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "cvtx_oracle.h"

/* ---- interpolants, reference src/RedistFunc.cpp:36-96 ---------------------------------- */
static const float k_radius[5] = {0.5f, 1.0f, 1.5f, 2.0f, 2.0f};

float cvtx_oracle_redist(int which, float U) {
	switch (which) {
	case CVTX_ORACLE_LAMBDA0: return U < 0.5f ? 1.f : 0.f;                                           /* :36-39 */
	case CVTX_ORACLE_LAMBDA1: return U <= 1.f ? 1.f - U : 0.f;                                       /* :48-51 */
	case CVTX_ORACLE_LAMBDA2:                                                                          /* :60-63 */
		return U < 0.5f ? 1.f - U * U : (U < 1.5f ? 0.5f * (1.f - U) * (2.f - U) : 0.f);
	case CVTX_ORACLE_LAMBDA3:                                                                          /* :72-76 */
		return U < 1.f ? 0.5f * (1.f - U * U) * (2.f - U)
		               : (U < 2.f ? (1.f / 6.f) * (1.f - U) * (2.f - U) * (3.f - U) : 0.f);
	default:                                                                                           /* M4', :85-89 */
		return U < 1.f ? 1.f - 2.5f * U * U + 1.5f * U * U * U
		               : (U < 2.f ? 0.5f * (1.f - U) * (2.f - U) * (2.f - U) : 0.f);
	}
}
float cvtx_oracle_redist_radius(int which) { return k_radius[which]; }

/* ---- a growing list of (node index, share) records ------------------------------------- */
typedef struct { uint32_t k[3]; uint64_t seq; float s[3]; } rec_t;
typedef struct { rec_t *v; size_t n, cap; } recs_t;

static void push(recs_t *r, const uint32_t *k, int dim, const float *s, int comps) {
	if (r->n == r->cap) {
		r->cap = r->cap ? 2 * r->cap : 4096;
		r->v = (rec_t *)realloc(r->v, r->cap * sizeof(rec_t));
	}
	rec_t *e = &r->v[r->n];
	memset(e, 0, sizeof(*e));
	for (int a = 0; a < dim; ++a) e->k[a] = k[a];
	for (int c = 0; c < comps; ++c) e->s[c] = s[c];
	e->seq = r->n++;
}

/* Depth-first order of the reference's trees: at every bit level, from the top, children are
 * visited in the order x + 2y + 4z of that bit (src/GridParticleOcttree.cpp:137-180,
 * src/UIntKey96.h:209-219).  Grid indices here never reach bit 31 (whose level is walked in
 * reverse there), so this is plain Morton order.  Ties fall back to arrival order. */
static int by_tree_order(const void *pa, const void *pb) {
	const rec_t *a = (const rec_t *)pa, *b = (const rec_t *)pb;
	for (int bit = 31; bit >= 0; --bit) {
		int la = 0, lb = 0;
		for (int ax = 0; ax < 3; ++ax) {
			la |= (int)((a->k[ax] >> bit) & 1u) << ax;
			lb |= (int)((b->k[ax] >> bit) & 1u) << ax;
		}
		if (la != lb) return la < lb ? -1 : 1;
	}
	return a->seq < b->seq ? -1 : (a->seq > b->seq ? 1 : 0);
}

/* ---- threshold search, reference src/redistribution_helper_funcs.cpp:32-91 -------------- */
static float strength_threshold(const float *strs, int n, int wanted) {
	enum { G = 1024 };
	float fminv = n > 0 ? strs[0] : 0.f, fmaxv = fminv, guesses[G];
	int counts[G], k = 0, interp;
	for (int i = 0; i < n; ++i) {                                  /* farray_info, :93-121 */
		fminv = fminv < strs[i] ? fminv : strs[i];
		fmaxv = fmaxv > strs[i] ? fmaxv : strs[i];
	}
	double minv = fminv, maxv = fmaxv, range;
	if (n < wanted) return (float)(maxv * 1.05);                   /* :43-45 */
	for (;;) {
		range = (maxv - minv) * 1.05;                              /* :51 */
		for (int i = 0; i < G; ++i) {
			counts[i] = 0;
			guesses[i] = (float)(minv + i * range / (float)(G - 1));   /* :54 */
		}
		for (int i = 0; i < n; ++i) {                              /* :56-61 */
			interp = (int)floor((double)(G - 1) * (strs[i] - minv) / range);
			if (interp < 0) counts[0]++;
			else if (interp >= G) counts[G - 1]++;
			else counts[interp]++;
		}
		interp = counts[G - 1];                                    /* :62-72 */
		for (int i = G - 2; i >= 0; --i) {
			maxv = guesses[i + 1];
			minv = guesses[i];
			interp += counts[i];
			counts[i] = interp;
			if (interp > wanted) { k = i + 1; break; }
		}
		if (minv == maxv || counts[k] == counts[k - 1] ||            /* :74-78 */
		    fabs((float)(wanted - counts[k]) / ((float)wanted)) < 0.01f * 0.6)
			break;
	}
	return guesses[k];                                             /* :80 */
}

/* ---- pruning, reference src/P3D.cpp:636-665 / src/P2D.cpp:407-436 ----------------------- */
static int remove_weak(float *pos, float *w, const float *strs, int n, int dim, int comps, float min_keep, int max_keepable) {
	float deficit[3] = {0.f, 0.f, 0.f};
	int j = 0;
	for (int i = 0; i < n; ++i) {
		if (strs[i] > min_keep && i < max_keepable) {
			for (int a = 0; a < dim; ++a) pos[j * dim + a] = pos[i * dim + a];
			for (int c = 0; c < comps; ++c) w[j * comps + c] = w[i * comps + c];
			++j;
		} else {
			for (int c = 0; c < comps; ++c) deficit[c] = w[i * comps + c] + deficit[c];
		}
	}
	for (int c = 0; c < comps; ++c) deficit[c] = deficit[c] / (float)j;
	for (int i = 0; i < j; ++i)
		for (int c = 0; c < comps; ++c) w[i * comps + c] = w[i * comps + c] + deficit[c];
	return j;
}

static void strengths_of(const float *w, int n, int comps, float *strs) {
	for (int i = 0; i < n; ++i)
		strs[i] = comps == 1 ? fabsf(w[i])
		                     : sqrtf(w[3 * i] * w[3 * i] + w[3 * i + 1] * w[3 * i + 1] + w[3 * i + 2] * w[3 * i + 2]);
}

/* ---- the redistribution itself, reference src/P3D.cpp:509-634 / src/P2D.cpp:283-405 ------ */
static int redistribute(int dim, const float *rows, int n, float *out, int max_out, int have_out, int which, float h, float negligible) {
	const int row = dim == 3 ? 7 : 4, comps = dim == 3 ? 3 : 1;
	if (n <= 0) return 0;
	const float rh = 1.f / h;                                      /* P3D.cpp:527 */
	const int R = (int)roundf(k_radius[which]);                    /* :542 */
	float lo[3], origin[3] = {0.f, 0.f, 0.f};
	double sum[3] = {0., 0., 0.};
	for (int a = 0; a < dim; ++a) lo[a] = rows[a];
	for (int i = 0; i < n; ++i)                                     /* minmax_xyz_posn, mean_xyz_posn */
		for (int a = 0; a < dim; ++a) {
			const float x = rows[i * row + a];
			lo[a] = lo[a] > x ? x : lo[a];
			sum[a] += (double)x;
		}
	for (int a = 0; a < dim; ++a) {                                /* :543-549 */
		const float mean = (float)(sum[a] / (double)n);
		const float corner = lo[a] - 1.f * (R * h);
		float d = (mean - corner) / h;
		d = roundf(d) + 5;
		origin[a] = mean - d * h;
	}

	recs_t recs = {0, 0, 0};
	for (int i = 0; i < n; ++i) {                                   /* :565-584 */
		const float *p = rows + i * row;
		uint32_t k0[3] = {0, 0, 0};
		for (int a = 0; a < dim; ++a) {
			if (dim == 3) {                                         /* UIntKey96.cpp:99-106 */
				k0[a] = (unsigned int)roundf((p[a] - origin[a]) * rh);
			} else {                                                /* UIntKey64.cpp:86-94 */
				double t = (double)p[a];
				t = (t - origin[a]) * rh;
				k0[a] = (unsigned int)roundf(t);
			}
		}
		for (int ii = -R; ii <= R; ++ii)                            /* nearby_keys, UIntKey96.cpp:52-65 */
			for (int jj = -R; jj <= R; ++jj)
				for (int kk = (dim == 3 ? -R : 0); kk <= (dim == 3 ? R : 0); ++kk) {
					uint32_t k[3] = {k0[0] + (uint32_t)ii, k0[1] + (uint32_t)jj, k0[2] + (uint32_t)kk};
					float f = 1.f, s[3] = {0.f, 0.f, 0.f};
					for (int a = 0; a < dim; ++a) {                 /* :573-581 */
						float node = origin[a];
						node += h * k[a];                           /* to_position_min */
						const float U = fabsf((p[a] - node) * rh);
						f = a == 0 ? cvtx_oracle_redist(which, U) : f * cvtx_oracle_redist(which, U);
					}
					int any = 0;
					for (int c = 0; c < comps; ++c) { s[c] = p[dim + c] * f; any |= s[c] != 0.f; }
					if (any) push(&recs, k, dim, s, comps);         /* zero shares are not inserted */
				}
	}
	qsort(recs.v, recs.n, sizeof(rec_t), by_tree_order);

	/* nodes: FP32 sums in arrival order, as one tree would form them */
	size_t cap = recs.n ? recs.n : 1;
	float *pos = (float *)calloc(cap * dim, sizeof(float)), *w = (float *)calloc(cap * comps, sizeof(float));
	int m = 0;
	for (size_t j = 0; j < recs.n;) {
		size_t e = j;
		float acc[3] = {recs.v[j].s[0], recs.v[j].s[1], recs.v[j].s[2]};
		for (e = j + 1; e < recs.n && !memcmp(recs.v[e].k, recs.v[j].k, sizeof(recs.v[j].k)); ++e)
			for (int c = 0; c < comps; ++c) acc[c] = acc[c] + recs.v[e].s[c];
		for (int a = 0; a < dim; ++a) { float x = origin[a]; x += h * recs.v[j].k[a]; pos[m * dim + a] = x; }
		for (int c = 0; c < comps; ++c) w[m * comps + c] = acc[c];
		++m;
		j = e;
	}
	free(recs.v);

	float *strs = (float *)malloc((size_t)(m ? m : 1) * sizeof(float));
	strengths_of(w, m, comps, strs);                               /* :604-607 */
	float ave = 0.f;
	for (int i = 0; i < m; ++i) ave += strs[i];                    /* farray_info(.., &mean, NULL, NULL) */
	ave /= m;
	float min_keep = ave * negligible;                             /* :609 */
	m = remove_weak(pos, w, strs, m, dim, comps, min_keep, m);     /* :610-612 */
	if (have_out) {
		if (m > max_out) {                                          /* :621-626 */
			strengths_of(w, m, comps, strs);
			min_keep = strength_threshold(strs, m, max_out);
			m = remove_weak(pos, w, strs, m, dim, comps, min_keep, max_out);
		}
		const float size = dim == 3 ? h * h * h : h * h;           /* :598 */
		for (int i = 0; i < m; ++i) {
			for (int a = 0; a < dim; ++a) out[i * row + a] = pos[i * dim + a];
			for (int c = 0; c < comps; ++c) out[i * row + dim + c] = w[i * comps + c];
			out[i * row + dim + comps] = size;
		}
	}
	free(pos); free(w); free(strs);
	return m;
}

int cvtx_oracle_P3D_redistribute(const float *src7, int n, float *out7, int max_out, int have_out, int which, float h, float negligible) {
	return redistribute(3, src7, n, out7, max_out, have_out, which, h, negligible);
}
int cvtx_oracle_P2D_redistribute(const float *src4, int n, float *out4, int max_out, int have_out, int which, float h, float negligible) {
	return redistribute(2, src4, n, out4, max_out, have_out, which, h, negligible);
}

/* ---- Pedrizzetti relaxation, reference src/P3D.cpp:667-707 ------------------------------- */
void cvtx_oracle_P3D_pedrizzetti(const float *src7, int n, float fdt, int reg, float sigma, float *out7) {
	if (n <= 0) return;
	float *pts = (float *)malloc((size_t)n * 3 * sizeof(float)), *om = (float *)malloc((size_t)n * 3 * sizeof(float));
	for (int i = 0; i < n; ++i)
		for (int a = 0; a < 3; ++a) pts[3 * i + a] = src7[7 * i + a];
	cvtx_oracle_P3D_M2M_vort_f32(src7, n, pts, n, om, reg, sigma);  /* :684-685 */
	const float tmp = 1.f - fdt;                                    /* :687 */
	memcpy(out7, src7, (size_t)n * 7 * sizeof(float));
	for (int i = 0; i < n; ++i) {                                   /* :689-699 */
		const float *ov = src7 + 7 * i + 3, *w = om + 3 * i;
		const float absomega = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
		const float coeff = sqrtf(ov[0] * ov[0] + ov[1] * ov[1] + ov[2] * ov[2]) / absomega;
		for (int a = 0; a < 3; ++a) {
			const float nv = ov[a] * tmp + w[a] * (coeff * fdt);
			out7[7 * i + 3 + a] = absomega != 0.f ? nv : 0.f;
		}
	}
	free(pts); free(om);
}
