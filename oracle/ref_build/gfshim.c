/* TEST INFRASTRUCTURE ONLY (oracle/_ref build glue) -- not product code.
 *
 * Stand-in for libgfortran.so.3 (GCC-5 Fortran runtime ABI).  The reference ships
 * prebuilt Ipopt 3.12.7 / MUMPS 4.10.0 binaries
 *   /root/reference/Ipopt-3.12.7/ThirdParty/Mumps/.libs/libcoinmumps.so.1.6.0
 * which import 22 versioned `_gfortran_*` runtime symbols; this image has no gfortran.
 * MUMPS only uses them for (silenced) diagnostics printing, string handling of option
 * names and packing of array sections, so small C equivalents are enough to load and
 * run the unmodified reference binaries as the parity oracle.
 *
 * GCC-5 ABI facts used: gfc_charlen_type = int; array descriptor =
 *   { void* base; size_t offset; ssize_t dtype; { ssize_t stride, lbound, ubound; } dim[7]; }
 *   rank = dtype & 7, element size = dtype >> 6.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

typedef struct { ssize_t stride, lbound, ubound; } gf_dim;
typedef struct { char* base; size_t offset; ssize_t dtype; gf_dim dim[7]; } gf_desc;

/* ---- I/O: MUMPS output is switched off by Ipopt (ICNTL(1..4)=0); swallow everything ---- */
void _gfortran_st_open(void* p) { (void)p; }
void _gfortran_st_close(void* p) { (void)p; }
void _gfortran_st_write(void* p) { (void)p; }
void _gfortran_st_write_done(void* p) { (void)p; }
void _gfortran_transfer_array_write(void* a, void* b, int c, int d) { (void)a; (void)b; (void)c; (void)d; }
void _gfortran_transfer_character_write(void* a, void* b, int c) { (void)a; (void)b; (void)c; }
void _gfortran_transfer_integer_write(void* a, void* b, int c) { (void)a; (void)b; (void)c; }
void _gfortran_transfer_logical_write(void* a, void* b, int c) { (void)a; (void)b; (void)c; }
void _gfortran_transfer_real_write(void* a, void* b, int c) { (void)a; (void)b; (void)c; }

/* ---- fatal paths ---- */
void _gfortran_os_error(const char* m) { fprintf(stderr, "gfshim: os_error: %s\n", m); abort(); }
void _gfortran_runtime_error(const char* m, ...) { fprintf(stderr, "gfshim: runtime_error: %s\n", m); abort(); }
void _gfortran_runtime_error_at(const char* w, const char* m, ...) {
  fprintf(stderr, "gfshim: runtime_error_at %s: %s\n", w, m); abort();
}
void _gfortran_stop_string(const char* s, int l) {
  if (s && l > 0) fprintf(stderr, "gfshim: STOP %.*s\n", l, s);
  exit(0);
}

/* ---- CHARACTER intrinsics ---- */
int _gfortran_string_len_trim(int len, const char* s) {
  while (len > 0 && s[len - 1] == ' ') --len;
  return len;
}
void _gfortran_string_trim(int* len, char** dst, int slen, const char* src) {
  int l = _gfortran_string_len_trim(slen, src);
  *len = l;
  *dst = NULL;
  if (l > 0) { *dst = (char*)malloc((size_t)l); memcpy(*dst, src, (size_t)l); }
}
void _gfortran_adjustl(char* dst, int len, const char* src) {
  int lead = 0;
  while (lead < len && src[lead] == ' ') ++lead;
  memmove(dst, src + lead, (size_t)(len - lead));
  memset(dst + (len - lead), ' ', (size_t)lead);
}
void _gfortran_concat_string(int dlen, char* dst, int l1, const char* s1, int l2, const char* s2) {
  int n1 = l1 < dlen ? l1 : dlen;
  memcpy(dst, s1, (size_t)n1);
  int rest = dlen - n1;
  int n2 = l2 < rest ? l2 : rest;
  if (n2 > 0) memcpy(dst + n1, s2, (size_t)n2);
  if (rest - n2 > 0) memset(dst + n1 + n2, ' ', (size_t)(rest - n2));
}
/* Fortran string comparison: shorter operand is blank padded. */
static int fcompare(const char* a, int la, const char* b, int lb) {
  int n = la < lb ? la : lb;
  int r = memcmp(a, b, (size_t)n);
  if (r) return r;
  for (int i = n; i < la; ++i) if (a[i] != ' ') return (unsigned char)a[i] > ' ' ? 1 : -1;
  for (int i = n; i < lb; ++i) if (b[i] != ' ') return (unsigned char)b[i] > ' ' ? -1 : 1;
  return 0;
}
typedef struct { char* low; int low_len; char* high; int high_len; int address; } gf_select_case;
int _gfortran_select_string(gf_select_case* table, int ncases, const char* sel, int sel_len) {
  int dflt = -1;
  for (int i = 0; i < ncases; ++i) {
    gf_select_case* c = &table[i];
    if (!c->low && !c->high) { dflt = c->address; continue; }
    if (c->low && fcompare(sel, sel_len, c->low, c->low_len) < 0) continue;
    if (c->high && fcompare(sel, sel_len, c->high, c->high_len) > 0) continue;
    return c->address;
  }
  return dflt;
}

/* ---- array sections ---- */
ssize_t _gfortran_size0(gf_desc* a) {
  int rank = (int)(a->dtype & 7);
  ssize_t total = 1;
  for (int n = 0; n < rank; ++n) {
    ssize_t e = a->dim[n].ubound - a->dim[n].lbound + 1;
    total *= e > 0 ? e : 0;
  }
  return total;
}
static int section_extents(gf_desc* a, ssize_t* ext, ssize_t* total, int* contiguous) {
  int rank = (int)(a->dtype & 7);
  ssize_t run = 1;
  *contiguous = 1;
  for (int n = 0; n < rank; ++n) {
    ext[n] = a->dim[n].ubound - a->dim[n].lbound + 1;
    if (ext[n] <= 0) return -1;
    if (a->dim[n].stride != run) *contiguous = 0;
    run *= ext[n];
  }
  *total = run;
  return rank;
}
static void section_copy(gf_desc* a, int rank, const ssize_t* ext, char* packed, int gather) {
  ssize_t esz = a->dtype >> 6, idx[7] = {0};
  for (;;) {
    ssize_t off = 0;
    for (int n = 0; n < rank; ++n) off += idx[n] * a->dim[n].stride;
    if (gather) memcpy(packed, a->base + off * esz, (size_t)esz);
    else memcpy(a->base + off * esz, packed, (size_t)esz);
    packed += esz;
    int n = 0;
    while (n < rank && ++idx[n] == ext[n]) idx[n++] = 0;
    if (n == rank) break;
  }
}
void* _gfortran_internal_pack(gf_desc* a) {
  ssize_t ext[7], total; int contiguous;
  int rank = section_extents(a, ext, &total, &contiguous);
  if (rank <= 0 || contiguous) return a->base;
  char* buf = (char*)malloc((size_t)(total * (a->dtype >> 6)));
  section_copy(a, rank, ext, buf, 1);
  return buf;
}
void _gfortran_internal_unpack(gf_desc* a, const void* src) {
  if (!src || src == (const void*)a->base) return;
  ssize_t ext[7], total; int contiguous;
  int rank = section_extents(a, ext, &total, &contiguous);
  if (rank <= 0) return;
  section_copy(a, rank, ext, (char*)src, 0);
}

void _gfortran_random_r8(double* x) { *x = drand48(); }
