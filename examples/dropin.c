/* examples/dropin.c -- a C caller written against cvortex's public header only.
 *
 *   gcc -std=c99 -Iinclude examples/dropin.c -Lcvortex_b200/lib -lcvortex -lm \
 *       -Wl,-rpath,$PWD/cvortex_b200/lib -o dropin && ./dropin [n]
 *
 * Nothing in here knows about CUDA: it is the calling sequence of the reference's README
 * (cvtx_initialise, build particles, cvtx_P3D_M2M_vel, cvtx_finalise).  It runs the same
 * call twice, with the accelerators enabled and disabled (the reference's documented CPU/GPU
 * switch), and prints how the two results compare.  Exit code 0 when they agree to 1e-5
 * relative L2.
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
	cvtx_finalise();
	free(particles); free(pparticles); free(mes); free(gpu); free(cpu);
	return rel <= 1e-5 ? 0 : 1;
}
