/* examples/dropin.c -- a C caller written against cvortex's public header only.
 *
 *   gcc -std=c99 -Iinclude examples/dropin.c -Lcvortex_b200/lib -lcvortex -lm \
 *       -Wl,-rpath,$PWD/cvortex_b200/lib -o dropin && ./dropin [n]
 *
 * Nothing in here knows about CUDA: it is the calling sequence of the reference's README
 * (cvtx_initialise, build particles, cvtx_P3D_M2M_vel, cvtx_finalise).  It runs the same
 * call twice, with the accelerators enabled and disabled (the reference's documented CPU/GPU
 * switch), and prints how the two results compare; then it remeshes the particles onto a grid
 * with M4' the same two ways (cvtx_P3D_redistribute_on_grid: ask for the count with a NULL
 * array, then for the particles).  Exit code 0 when the velocities agree to 1e-5 relative L2
 * and the two redistributions are the same particles.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <cvortex/libcvtx.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static float urand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return 10.0f * (float)(*s >> 8) / 16777216.0f; }

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 20000;
	unsigned seed = 12345u;
	cvtx_P3D *particles = malloc(sizeof(cvtx_P3D) * n);
	const cvtx_P3D **pparticles = malloc(sizeof(cvtx_P3D *) * n);
	bsv_V3f *mes = malloc(sizeof(bsv_V3f) * n), *gpu = malloc(sizeof(bsv_V3f) * n), *cpu = malloc(sizeof(bsv_V3f) * n);
	for (int i = 0; i < n; ++i) {
		for (int c = 0; c < 3; ++c) {
			particles[i].coord.x[c] = urand(&seed);
			particles[i].vorticity.x[c] = urand(&seed);
			mes[i].x[c] = urand(&seed);
		}
		particles[i].volume = 0.01f;
		pparticles[i] = &particles[i];
	}

	cvtx_initialise();
	printf("%s", cvtx_information());
	const int n_acc = cvtx_num_accelerators();
	cvtx_VortFunc vf = cvtx_VortFunc_winckelmans();

	double t0 = now();
	cvtx_P3D_M2M_vel(pparticles, n, mes, n, gpu, &vf, 0.02f);           /* accelerator 0 is on by default */
	double t_on = now() - t0;
	for (int k = 0; k < n_acc; ++k) cvtx_accelerator_disable(k);        /* the reference's CPU switch */
	t0 = now();
	cvtx_P3D_M2M_vel(pparticles, n, mes, n, cpu, &vf, 0.02f);
	double t_off = now() - t0;
	for (int k = 0; k < n_acc; ++k) cvtx_accelerator_enable(k);

	double num = 0, den = 0;
	for (int i = 0; i < n; ++i)
		for (int c = 0; c < 3; ++c) {
			const double d = (double)gpu[i].x[c] - cpu[i].x[c];
			num += d * d; den += (double)cpu[i].x[c] * cpu[i].x[c];
		}
	const double rel = sqrt(num / den);
	printf("%d x %d pairs: accelerators on %.3f ms, off %.3f ms (%d accelerator%s), relative L2 difference %.2e\n",
	       n, n, 1e3 * t_on, 1e3 * t_off, n_acc, n_acc == 1 ? "" : "s", rel);

	/* ---- redistribution onto a grid, as the reference's README describes it ---- */
	const cvtx_RedistFunc m4p = cvtx_RedistFunc_m4p();
	const float h = 10.0f * (float)cbrt(2.0 / n);                          /* about two particles per cell */
	t0 = now();
	const int count = cvtx_P3D_redistribute_on_grid(pparticles, n, NULL, 0, &m4p, h, 1e-4f);
	cvtx_P3D *fresh = malloc(sizeof(cvtx_P3D) * (count > 0 ? count : 1)), *fresh_host = malloc(sizeof(cvtx_P3D) * (count > 0 ? count : 1));
	const int made = cvtx_P3D_redistribute_on_grid(pparticles, n, fresh, count, &m4p, h, 1e-4f);
	const double t_remesh = now() - t0;
	for (int k = 0; k < n_acc; ++k) cvtx_accelerator_disable(k);
	const int made_host = cvtx_P3D_redistribute_on_grid(pparticles, n, fresh_host, count, &m4p, h, 1e-4f);
	for (int k = 0; k < n_acc; ++k) cvtx_accelerator_enable(k);
	int same = made == count && made_host == count;
	for (int i = 0; same && i < made; ++i)
		for (int c = 0; c < 3; ++c)
			same = fresh[i].coord.x[c] == fresh_host[i].coord.x[c] && fresh[i].vorticity.x[c] == fresh_host[i].vorticity.x[c];
	double in[3] = {0, 0, 0}, out[3] = {0, 0, 0};
	for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) in[c] += particles[i].vorticity.x[c];
	for (int i = 0; i < made; ++i) for (int c = 0; c < 3; ++c) out[c] += fresh[i].vorticity.x[c];
	printf("redistribution (M4', h = %.3f): %d -> %d particles in %.3f ms (count + fill), total vorticity %.6g -> %.6g, "
	       "accelerators on / off give %s particles\n", h, n, made, 1e3 * t_remesh, in[0], out[0], same ? "the same" : "DIFFERENT");

	cvtx_finalise();
	free(particles); free(pparticles); free(mes); free(gpu); free(cpu); free(fresh); free(fresh_host);
	return rel <= 1e-5 && same ? 0 : 1;
}
