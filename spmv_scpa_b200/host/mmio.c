/* mmio.c -- banner and size-line reader/writer for Matrix Market files.
 *
 * Fresh implementation of the subset declared in include/mmio.h.  Accept /
 * reject behaviour follows the NIST routines vendored by the reference
 * (reference src/mmio.c:93-166 mm_read_banner, :175-200
 * mm_read_mtx_crd_size): five whitespace separated banner tokens, the first
 * must start with "%%MatrixMarket" (case sensitive), the other four are
 * matched case-insensitively.
 */
#include <ctype.h>
#include <string.h>

#include "mmio.h"

struct keyword {
      const char *word;
      char code;
};

static const struct keyword k_format[] = {{"coordinate", 'C'}, {"array", 'A'}};
static const struct keyword k_field[] = {
    {"real", 'R'}, {"complex", 'C'}, {"pattern", 'P'}, {"integer", 'I'}};
static const struct keyword k_symmetry[] = {{"general", 'G'},
                                            {"symmetric", 'S'},
                                            {"hermitian", 'H'},
                                            {"skew-symmetric", 'K'}};

static void lower_in_place(char *s) {
      for (; *s; ++s)
            *s = (char)tolower((unsigned char)*s);
}

static int lookup(const struct keyword *tab, size_t n, const char *tok,
                  char *out) {
      for (size_t i = 0; i < n; ++i) {
            if (strcmp(tab[i].word, tok) == 0) {
                  *out = tab[i].code;
                  return 0;
            }
      }
      return -1;
}

#define LOOKUP(tab, tok, out) lookup(tab, sizeof(tab) / sizeof((tab)[0]), tok, out)

int mm_read_banner(FILE *f, MM_typecode *matcode) {
      char line[MM_MAX_LINE_LENGTH];
      char tok[5][MM_MAX_TOKEN_LENGTH];

      mm_clear_typecode(matcode);

      if (!fgets(line, sizeof line, f))
            return MM_PREMATURE_EOF;
      if (sscanf(line, "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2],
                 tok[3], tok[4]) != 5)
            return MM_PREMATURE_EOF;

      if (strncmp(tok[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0)
            return MM_NO_HEADER;

      for (int i = 1; i < 5; ++i)
            lower_in_place(tok[i]);

      if (strcmp(tok[1], "matrix") != 0)
            return MM_UNSUPPORTED_TYPE;
      (*matcode)[0] = 'M';

      if (LOOKUP(k_format, tok[2], &(*matcode)[1]) ||
          LOOKUP(k_field, tok[3], &(*matcode)[2]) ||
          LOOKUP(k_symmetry, tok[4], &(*matcode)[3]))
            return MM_UNSUPPORTED_TYPE;

      return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz) {
      char line[MM_MAX_LINE_LENGTH];

      *M = *N = *nz = 0;

      /* comment lines start with '%' in column 0 */
      do {
            if (!fgets(line, sizeof line, f))
                  return MM_PREMATURE_EOF;
      } while (line[0] == '%');

      if (sscanf(line, "%d %d %d", M, N, nz) == 3)
            return 0;

      /* The first non-comment line was blank (or partial): the three sizes
       * may follow on later lines.  The NIST loop retries forever on a
       * non-numeric token; here that case is an error. */
      int got = fscanf(f, "%d %d %d", M, N, nz);
      if (got == 3)
            return 0;
      *M = *N = *nz = 0;
      return MM_PREMATURE_EOF;
}

static const char *word_of(const struct keyword *tab, size_t n, char code) {
      for (size_t i = 0; i < n; ++i)
            if (tab[i].code == code)
                  return tab[i].word;
      return NULL;
}

int mm_write_banner(FILE *f, MM_typecode matcode) {
      const char *fmt = word_of(k_format, 2, matcode[1]);
      const char *fld = word_of(k_field, 4, matcode[2]);
      const char *sym = word_of(k_symmetry, 4, matcode[3]);
      if (matcode[0] != 'M' || !fmt || !fld || !sym)
            return MM_UNSUPPORTED_TYPE;
      if (fprintf(f, "%s matrix %s %s %s\n", MatrixMarketBanner, fmt, fld,
                  sym) < 0)
            return MM_COULD_NOT_WRITE_FILE;
      return 0;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz) {
      return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? MM_COULD_NOT_WRITE_FILE
                                                    : 0;
}
