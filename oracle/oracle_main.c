/*
 * oracle_main.c -- command-line twin of the reference's `compareLbm <paramfile> <obstaclefile>`
 * (main/LastChance.cpp:120-288) built on the restated oracle.  TEST INFRASTRUCTURE ONLY.
 *
 *   oracle_lbm <paramfile> <obstaclefile> [--double] [--threads N] [--iters N] [--outdir DIR]
 *              [--transposed-flag] [--no-final-state]
 *
 * Writes av_vels.dat and final_state.dat (reference C format) into DIR (default: cwd) and prints the
 * same "==done==" / "Reynolds number:" / "Elapsed time:" lines as LastChance.cpp:276-279.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "lbm_oracle.h"

static double now_s(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + tv.tv_usec / 1000000.0;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "Usage: %s <paramfile> <obstaclefile> [--double] [--threads N] [--iters N] [--outdir DIR]\n",
            argv[0]);
    return EXIT_FAILURE;
  }
  int accum = ORACLE_ACCUM_FLOAT, threads = 1, iters = -1, transposed = 0, write_state = 1;
  const char *outdir = ".";
  for (int i = 3; i < argc; ++i) {
    if (!strcmp(argv[i], "--double")) accum = ORACLE_ACCUM_DOUBLE;
    else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--outdir") && i + 1 < argc) outdir = argv[++i];
    else if (!strcmp(argv[i], "--transposed-flag")) transposed = 1;
    else if (!strcmp(argv[i], "--no-final-state")) write_state = 0;
    else { fprintf(stderr, "unknown option %s\n", argv[i]); return EXIT_FAILURE; }
  }
  oracle_params p;
  int rc = oracle_load_params(argv[1], &p);
  if (rc) { fprintf(stderr, "could not read param file %s (code %d)\n", argv[1], rc); return EXIT_FAILURE; }
  if (iters < 0) iters = p.max_iters;
  const size_t n = (size_t)p.nx * (size_t)p.ny;
  float *cells = malloc(n * 9 * sizeof(float)), *tmp = malloc(n * 9 * sizeof(float));
  float *av = malloc(sizeof(float) * (size_t)(iters > 0 ? iters : 1));
  uint8_t *obst = malloc(n);
  if (!cells || !tmp || !av || !obst) { fprintf(stderr, "cannot allocate memory\n"); return EXIT_FAILURE; }
  long lines = oracle_load_obstacles(argv[2], p.nx, p.ny, obst);
  if (lines < 0) { fprintf(stderr, "could not read obstacle file %s (code %ld)\n", argv[2], lines); return EXIT_FAILURE; }
  oracle_init_cells(&p, cells);

  const double tic = now_s();
  float *final = oracle_run(&p, cells, tmp, obst, iters, av, accum, threads, 1);
  const double toc = now_s();

  printf("==done==\n");
  printf("Reynolds number:\t\t%.12E\n", oracle_reynolds(&p, oracle_av_velocity(&p, final, obst)));
  printf("Elapsed time:\t\t\t%.6lf (s)\n", toc - tic);
  printf("MLUPS:\t\t\t\t%.3f\n", (double)n * iters / (toc - tic) / 1e6);

  char path[4096];
  if (write_state) {
    snprintf(path, sizeof path, "%s/final_state.dat", outdir);
    if (oracle_write_final_state(path, &p, final, obst, transposed)) return EXIT_FAILURE;
  }
  snprintf(path, sizeof path, "%s/av_vels.dat", outdir);
  if (oracle_write_av_vels(path, av, iters)) return EXIT_FAILURE;
  free(cells); free(tmp); free(av); free(obst);
  return EXIT_SUCCESS;
}
