/*
 * oracle_boundary.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 * Boundary-element flux AsBMFG -> e3b (compressible/asbmfg.f, e3b.f, e3bvar.f).
 */
#include "oracle_internal.h"
#include <stdio.h>
#include <stdlib.h>

void orc_tri_tables(int rule, int *nintb, double *Qwtb, double *shpb,
                    double *shglb) {
  (void)rule; (void)nintb; (void)Qwtb; (void)shpb; (void)shglb;
  fprintf(stderr, "orc_tri_tables: not restated yet\n");
  abort();
}

void orc_asbmfg(const orc_part *p, int iblk, double *res) {
  (void)p; (void)iblk; (void)res;
  fprintf(stderr, "orc_asbmfg: boundary elements not restated yet\n");
  abort();
}
