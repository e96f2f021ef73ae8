/* pomo_state.c -- allocation and name->array registry of the CPU ORACLE.
 * TEST INFRASTRUCTURE ONLY (see pomo.h).  Mirrors the COMMON blocks of
 * pom.h_dist:142-198,208-212,291-364,410-450,532-608 with run-time sizes. */
#include "pomo.h"
#include <stdlib.h>
#include <string.h>

static double *zalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

pomo_t *pomo_create(int im, int jm, int kb) {
  pomo_t *S = (pomo_t *)calloc(1, sizeof(pomo_t));
  S->im = im; S->jm = jm; S->kb = kb;
  S->imm1 = im - 1; S->imm2 = im - 2;
  S->jmm1 = jm - 1; S->jmm2 = jm - 2;
  S->kbm1 = kb - 1; S->kbm2 = kb - 2;
  size_t n3 = (size_t)im * jm * kb, n2 = (size_t)im * jm;
#define X(n) S->n = zalloc(n3);
  POMO_F3D(X)
#undef X
#define X(n) S->n = zalloc(n2);
  POMO_F2D(X)
#undef X
#define X(n) S->n = zalloc(jm);
  POMO_BJ(X)
#undef X
#define X(n) S->n = zalloc(im);
  POMO_BI(X)
#undef X
#define X(n) S->n = zalloc((size_t)jm * kb);
  POMO_BJK(X)
#undef X
#define X(n) S->n = zalloc((size_t)im * kb);
  POMO_BIK(X)
#undef X
#define X(n) S->n = zalloc(kb);
  POMO_F1D(X)
#undef X
  for (int i = 0; i < POMO_NSCR3; ++i) S->scr3[i] = zalloc(n3);
  for (int i = 0; i < POMO_NSCR2; ++i) S->scr2[i] = zalloc(n2);
  /* one sub-domain: all neighbours -1 (parallel_mpi.f:109-119) */
  S->n_west = S->n_east = S->n_south = S->n_north = -1;
  return S;
}

void pomo_destroy(pomo_t *S) {
  if (!S) return;
#define X(n) free(S->n);
  POMO_F3D(X) POMO_F2D(X) POMO_BJ(X) POMO_BI(X) POMO_BJK(X) POMO_BIK(X) POMO_F1D(X)
#undef X
  for (int i = 0; i < POMO_NSCR3; ++i) free(S->scr3[i]);
  for (int i = 0; i < POMO_NSCR2; ++i) free(S->scr2[i]);
  free(S);
}

double *pomo_field(pomo_t *S, const char *name, long *n) {
  size_t n3 = (size_t)S->im * S->jm * S->kb, n2 = (size_t)S->im * S->jm;
#define X(f) if (!strcmp(name, #f)) { if (n) *n = (long)n3; return S->f; }
  POMO_F3D(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = (long)n2; return S->f; }
  POMO_F2D(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = S->jm; return S->f; }
  POMO_BJ(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = S->im; return S->f; }
  POMO_BI(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = (long)S->jm * S->kb; return S->f; }
  POMO_BJK(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = (long)S->im * S->kb; return S->f; }
  POMO_BIK(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { if (n) *n = S->kb; return S->f; }
  POMO_F1D(X)
#undef X
  if (n) *n = 0;
  return NULL;
}

int pomo_set(pomo_t *S, const char *name, double v) {
#define X(f) if (!strcmp(name, #f)) { S->f = v; return 0; }
  POMO_SCAL_D(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { S->f = (int)v; return 0; }
  POMO_SCAL_I(X)
#undef X
  return -1;
}

double pomo_get(pomo_t *S, const char *name) {
#define X(f) if (!strcmp(name, #f)) return S->f;
  POMO_SCAL_D(X)
#undef X
#define X(f) if (!strcmp(name, #f)) return (double)S->f;
  POMO_SCAL_I(X)
#undef X
  return 0.0 / 0.0;
}
