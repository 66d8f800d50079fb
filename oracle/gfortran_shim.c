/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * A minimal stand-in for libgfortran.so.3 (the GCC 4.8 Fortran run-time library), just large enough to run the
 * prebuilt `sph` executables that the reference repository ships next to its example inputs
 * (example_problems/<problem>/sph, "GCC: (GNU) 4.8.5", built from the sources in the same directory with
 * `gfortran -O3`, see the Makefile there). No Fortran compiler or run-time exists in this image, so this shim is
 * what lets the UNMODIFIED reference binary run here and pin the oracle (oracle/make_reference_goldens.py).
 *
 * Only the 25 entry points the three binaries import are implemented (nm -D sph), with the GCC 4.8 parameter-block
 * layout (libgfortran/io/io.h of that release; offsets checked against the binary's call sites with objdump):
 *   list-directed READ from files, list-directed and (A)/(Fw.d)/(Iw.m) formatted WRITE, ADVANCE='NO', internal
 *   WRITE, OPEN/CLOSE by unit + file name, string helpers, STOP.
 * Deliberate difference from the real library: every REAL is written with all its significant digits
 * (%.17g for kind 8, %.9g for kind 4) whatever the format says -- the reference prints F16.8, which would hide
 * everything below 1e-8; the golden vectors need the full value. The numbers themselves are the binary's own.
 *
 * Build: see oracle/Makefile (target _ref/libgfortran.so.3, with the GFORTRAN_1.0 / GFORTRAN_1.4 version nodes
 * the binary asks for).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

/* ---- parameter blocks (GCC 4.8, x86-64) ------------------------------------------------------------ */
typedef struct {
  int32_t flags;
  int32_t unit;
  const char *filename;
  int32_t line;
  int32_t iomsg_len;
  char *iomsg;
  int32_t *iostat;
} st_common; /* 40 bytes */

typedef struct {
  st_common c;
  int32_t recl_in;
  int32_t file_len;
  char *file;
  char *status;
  int32_t status_len;
  /* access, form, ... follow; not used by the reference */
} st_open_t;

typedef struct {
  st_common c;
  /* nothing else is read */
} st_close_t;

typedef struct {
  st_common c;
  int64_t rec;
  int64_t *size, *iolength;
  void *internal_unit_desc;
  char *format;
  int32_t format_len;
  int32_t advance_len;
  char *advance;
  char *internal_unit;
  int32_t internal_unit_len;
  int32_t namelist_name_len;
  char *namelist_name;
  /* private area of the real library follows */
} st_dt;

enum {
  IOPARM_OPEN_HAS_FILE = 1 << 8,
  IOPARM_DT_LIST_FORMAT = 1 << 7,
  IOPARM_DT_HAS_FORMAT = 1 << 12,
  IOPARM_DT_HAS_ADVANCE = 1 << 13,
  IOPARM_DT_HAS_INTERNAL_UNIT = 1 << 14
};

typedef struct {
  ssize_t stride, lbound, ubound;
} gf_dim;
typedef struct {
  char *base;
  size_t offset;
  ssize_t dtype; /* rank: bits 0-2, type: bits 3-5, element size: bits 6.. */
  gf_dim dim[7];
} gf_desc;
enum { BT_INTEGER = 1, BT_LOGICAL = 2, BT_REAL = 3, BT_CHARACTER = 6 };

/* ---- units ---------------------------------------------------------------------------------------- */
#define MAXOPEN 64 /* units open at the same time; unit NUMBERS are arbitrary (the reference uses nnode + step) */
#define LINECAP 65536
typedef struct {
  int number, used;
  FILE *fp;
  char *line; /* current input record (list-directed READ) */
  int pos, have_line;
} unit_t;
static unit_t units[MAXOPEN];

static void die(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "gfortran_shim: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  exit(3);
}

static unit_t *find_unit(int u) {
  for (int k = 0; k < MAXOPEN; ++k)
    if (units[k].used && units[k].number == u) return &units[k];
  return NULL;
}
static unit_t *new_unit(int u) {
  for (int k = 0; k < MAXOPEN; ++k)
    if (!units[k].used) {
      char *keep = units[k].line;
      memset(&units[k], 0, sizeof units[k]);
      units[k].line = keep;
      units[k].used = 1;
      units[k].number = u;
      return &units[k];
    }
  die("more than %d units open", MAXOPEN);
  return NULL;
}
static unit_t *get_unit(int u, int for_write) {
  unit_t *p = find_unit(u);
  if (p) return p;
  p = new_unit(u);
  if (u == 6)
    p->fp = stdout;
  else if (u == 0)
    p->fp = stderr;
  else if (u == 5)
    p->fp = stdin;
  else { /* implicit open: fort.N */
    char name[32];
    snprintf(name, sizeof name, "fort.%d", u);
    p->fp = fopen(name, for_write ? "w+" : "r");
    if (!p->fp) die("cannot open %s", name);
  }
  return p;
}
static void close_unit(unit_t *p) {
  if (p->fp && p->fp != stdout && p->fp != stderr && p->fp != stdin) fclose(p->fp);
  p->fp = NULL;
  p->used = 0;
  p->have_line = 0;
}

void _gfortran_st_open(st_open_t *p) {
  const int u = p->c.unit;
  if (!(p->c.flags & IOPARM_OPEN_HAS_FILE)) die("OPEN without FILE= (unit %d, %s:%d)", u, p->c.filename, p->c.line);
  char name[4096];
  int n = p->file_len;
  while (n > 0 && p->file[n - 1] == ' ') --n; /* trailing blanks are not part of the name */
  if (n >= (int)sizeof name) die("file name too long");
  memcpy(name, p->file, n);
  name[n] = 0;
  unit_t *q = find_unit(u);
  if (q) close_unit(q);
  q = new_unit(u);
  q->fp = fopen(name, "r+"); /* STATUS='UNKNOWN': use the file if it exists, else create it */
  if (!q->fp) q->fp = fopen(name, "w+");
  if (!q->fp) die("cannot open '%s'", name);
}

void _gfortran_st_close(st_close_t *p) {
  unit_t *q = find_unit(p->c.unit);
  if (q) close_unit(q);
}

/* ---- current data-transfer statement (the reference never nests I/O) --------------------------------- */
static struct {
  int reading, list, advance_no, internal, nitems, fetched;
  unit_t *u;
  char fmt[256];
  char *ibuf; /* internal unit */
  int ilen, ipos;
} cur;

static void begin(st_dt *d, int reading) {
  memset(&cur, 0, sizeof cur);
  cur.reading = reading;
  cur.list = (d->c.flags & IOPARM_DT_LIST_FORMAT) != 0;
  if (d->c.flags & IOPARM_DT_HAS_FORMAT) {
    int n = d->format_len < 255 ? d->format_len : 255;
    memcpy(cur.fmt, d->format, n);
    cur.fmt[n] = 0;
  }
  if ((d->c.flags & IOPARM_DT_HAS_ADVANCE) && d->advance_len >= 1 && (d->advance[0] == 'n' || d->advance[0] == 'N'))
    cur.advance_no = 1;
  if (d->c.flags & IOPARM_DT_HAS_INTERNAL_UNIT) {
    cur.internal = 1;
    cur.ibuf = d->internal_unit;
    cur.ilen = d->internal_unit_len;
    cur.ipos = 0;
    if (!reading) memset(cur.ibuf, ' ', cur.ilen);
  } else {
    cur.u = get_unit(d->c.unit, !reading);
  }
}

/* ---- WRITE ------------------------------------------------------------------------------------------- */
static void out(const char *s, int n) {
  if (cur.internal) {
    for (int i = 0; i < n && cur.ipos < cur.ilen; ++i) cur.ibuf[cur.ipos++] = s[i];
  } else {
    fwrite(s, 1, n, cur.u->fp);
  }
}
static void outs(const char *s) { out(s, (int)strlen(s)); }

void _gfortran_st_write(st_dt *d) { begin(d, 0); }
void _gfortran_st_write_done(st_dt *d) {
  (void)d;
  if (!cur.internal && !cur.advance_no) outs("\n");
}

/* first edit descriptor of the format, e.g. "(I6.6)" -> 'I', w = 6, m = 6 */
static char fmt_desc(int *w, int *m) {
  const char *p = cur.fmt;
  *w = 0;
  *m = 0;
  while (*p && !isalpha((unsigned char)*p)) ++p;
  if (!*p) return 0;
  const char c = (char)toupper((unsigned char)*p++);
  *w = (int)strtol(p, (char **)&p, 10);
  if (*p == '.') *m = (int)strtol(p + 1, NULL, 10);
  return c;
}

static void put_sep(void) {
  if (cur.list) outs(" ");
  ++cur.nitems;
}
static void put_real(const void *p, int kind) {
  char b[64];
  put_sep();
  if (kind == 4)
    snprintf(b, sizeof b, "%.9g", (double)*(const float *)p);
  else if (kind == 8)
    snprintf(b, sizeof b, "%.17g", *(const double *)p);
  else
    die("REAL kind %d", kind);
  outs(b);
}
static void put_int(const void *p, int kind) {
  long long v = kind == 4 ? *(const int32_t *)p : kind == 8 ? *(const int64_t *)p : kind == 2 ? *(const int16_t *)p
                                                                                            : *(const int8_t *)p;
  char b[64];
  put_sep();
  int w, m;
  if (!cur.list && fmt_desc(&w, &m) == 'I' && w > 0) {
    char digits[48];
    if (m > 30) m = 30;
    snprintf(digits, sizeof digits, "%0*lld", m > 0 ? m : 1, v < 0 ? -v : v);
    int pad = w - (int)strlen(digits) - (v < 0 ? 1 : 0);
    while (pad-- > 0) outs(" ");
    if (v < 0) outs("-");
    outs(digits);
    return;
  }
  snprintf(b, sizeof b, "%lld", v);
  outs(b);
}
void _gfortran_transfer_real_write(st_dt *d, const void *p, int kind) {
  (void)d;
  put_real(p, kind);
}
void _gfortran_transfer_integer_write(st_dt *d, const void *p, int kind) {
  (void)d;
  put_int(p, kind);
}
void _gfortran_transfer_logical_write(st_dt *d, const void *p, int kind) {
  (void)d;
  const int v = kind == 4 ? *(const int32_t *)p != 0 : kind == 8 ? *(const int64_t *)p != 0 : *(const int8_t *)p != 0;
  put_sep();
  outs(v ? "T" : "F");
}
void _gfortran_transfer_character_write(st_dt *d, const void *p, int len) {
  (void)d;
  put_sep();
  out((const char *)p, len);
}
void _gfortran_transfer_array_write(st_dt *d, gf_desc *a, int kind, int charlen) {
  const int rank = (int)(a->dtype & 7), type = (int)((a->dtype >> 3) & 7);
  const ssize_t esize = a->dtype >> 6;
  ssize_t idx[7] = {0}, ext[7];
  ssize_t total = 1;
  for (int r = 0; r < rank; ++r) {
    ext[r] = a->dim[r].ubound - a->dim[r].lbound + 1;
    if (ext[r] <= 0) return;
    total *= ext[r];
  }
  for (ssize_t n = 0; n < total; ++n) {
    ssize_t off = 0;
    for (int r = 0; r < rank; ++r) off += idx[r] * a->dim[r].stride;
    const char *e = a->base + off * esize;
    if (type == BT_REAL)
      put_real(e, kind);
    else if (type == BT_INTEGER)
      put_int(e, kind);
    else if (type == BT_LOGICAL)
      _gfortran_transfer_logical_write(d, e, kind);
    else if (type == BT_CHARACTER)
      _gfortran_transfer_character_write(d, e, charlen);
    else
      die("array write of type %d", type);
    for (int r = 0; r < rank; ++r) {
      if (++idx[r] < ext[r]) break;
      idx[r] = 0;
    }
  }
}

/* ---- list-directed READ -------------------------------------------------------------------------------- */
void _gfortran_st_read(st_dt *d) {
  begin(d, 1);
  if (!cur.list) die("formatted READ is not supported (%s:%d)", d->c.filename, d->c.line);
  if (cur.internal) die("internal READ is not supported");
}
static int fetch_line(void) {
  unit_t *u = cur.u;
  if (!u->line) u->line = (char *)malloc(LINECAP);
  if (!fgets(u->line, LINECAP, u->fp)) return 0;
  u->pos = 0;
  u->have_line = 1;
  cur.fetched = 1;
  return 1;
}
/* next value of the record(s): returns its length (0 = null value), -1 at end of file */
static int token(char *buf, int cap) {
  unit_t *u = cur.u;
  for (;;) {
    if (!u->have_line && !fetch_line()) return -1;
    char *s = u->line;
    while (s[u->pos] == ' ' || s[u->pos] == '\t' || s[u->pos] == '\r') ++u->pos;
    if (s[u->pos] == '\n' || s[u->pos] == 0) {
      u->have_line = 0; /* record exhausted: the value is in a later record */
      continue;
    }
    break;
  }
  char *s = cur.u->line;
  int n = 0;
  int *pos = &cur.u->pos;
  if (s[*pos] == ',') { /* null value */
    ++*pos;
    buf[0] = 0;
    return 0;
  }
  if (s[*pos] == '\'' || s[*pos] == '"') {
    const char q = s[(*pos)++];
    while (s[*pos] && s[*pos] != '\n') {
      if (s[*pos] == q) {
        if (s[*pos + 1] == q) {
          ++*pos; /* doubled delimiter */
        } else {
          ++*pos;
          break;
        }
      }
      if (n < cap - 1) buf[n++] = s[*pos];
      ++*pos;
    }
  } else {
    while (s[*pos] && !strchr(" \t\r\n,/", s[*pos])) {
      if (n < cap - 1) buf[n++] = s[*pos];
      ++*pos;
    }
  }
  buf[n] = 0;
  while (s[*pos] == ' ' || s[*pos] == '\t' || s[*pos] == '\r') ++*pos; /* value separator: blanks, one comma */
  if (s[*pos] == ',') ++*pos;
  ++cur.nitems;
  return n;
}
void _gfortran_st_read_done(st_dt *d) {
  (void)d;
  if (!cur.fetched && !cur.u->have_line) fetch_line(); /* READ without items still consumes a record */
  cur.u->have_line = 0;                                /* the rest of the record is skipped */
}
static void eof(st_dt *d) { die("end of file in READ at %s:%d", d->c.filename, d->c.line); }

void _gfortran_transfer_real(st_dt *d, void *p, int kind) {
  if (!cur.reading) {
    put_real(p, kind);
    return;
  }
  char b[256];
  const int n = token(b, sizeof b);
  if (n < 0) eof(d);
  if (n == 0) return;
  for (char *c = b; *c; ++c)
    if (*c == 'd' || *c == 'D') *c = 'e';
  char *end;
  const double v = strtod(b, &end);
  if (end == b) die("bad real '%s' at %s:%d", b, d->c.filename, d->c.line);
  if (kind == 4)
    *(float *)p = (float)v; /* one rounding from the decimal string would be strtof; see below */
  else
    *(double *)p = v;
  if (kind == 4) *(float *)p = strtof(b, NULL);
}
void _gfortran_transfer_integer(st_dt *d, void *p, int kind) {
  if (!cur.reading) {
    put_int(p, kind);
    return;
  }
  char b[256];
  const int n = token(b, sizeof b);
  if (n < 0) eof(d);
  if (n == 0) return;
  char *end;
  const long long v = strtoll(b, &end, 10);
  if (end == b) die("bad integer '%s' at %s:%d", b, d->c.filename, d->c.line);
  if (kind == 4)
    *(int32_t *)p = (int32_t)v;
  else if (kind == 8)
    *(int64_t *)p = v;
  else
    die("INTEGER kind %d", kind);
}
void _gfortran_transfer_logical(st_dt *d, void *p, int kind) {
  if (!cur.reading) {
    _gfortran_transfer_logical_write(d, p, kind);
    return;
  }
  char b[256];
  const int n = token(b, sizeof b);
  if (n < 0) eof(d);
  if (n == 0) return;
  const char *c = b;
  if (*c == '.') ++c;
  int v;
  if (*c == 't' || *c == 'T')
    v = 1;
  else if (*c == 'f' || *c == 'F')
    v = 0;
  else
    die("bad logical '%s' at %s:%d", b, d->c.filename, d->c.line);
  if (kind == 4)
    *(int32_t *)p = v;
  else if (kind == 8)
    *(int64_t *)p = v;
  else
    *(int8_t *)p = (int8_t)v;
}
void _gfortran_transfer_character(st_dt *d, void *p, int len) {
  if (!cur.reading) {
    _gfortran_transfer_character_write(d, p, len);
    return;
  }
  char b[4096];
  const int n = token(b, sizeof b);
  if (n < 0) eof(d);
  if (n == 0) return;
  memset(p, ' ', len);
  memcpy(p, b, n < len ? n : len);
}

/* ---- strings, start-up, errors ------------------------------------------------------------------------- */
void _gfortran_concat_string(int destlen, char *dest, int len1, const char *s1, int len2, const char *s2) {
  if (len1 >= destlen) {
    memcpy(dest, s1, destlen);
    return;
  }
  memcpy(dest, s1, len1);
  dest += len1;
  destlen -= len1;
  if (len2 >= destlen) {
    memcpy(dest, s2, destlen);
    return;
  }
  memcpy(dest, s2, len2);
  memset(dest + len2, ' ', destlen - len2);
}
int _gfortran_string_len_trim(int len, const char *s) {
  while (len > 0 && s[len - 1] == ' ') --len;
  return len;
}
void _gfortran_string_trim(int *len, char **dest, int slen, const char *src) {
  while (slen > 0 && src[slen - 1] == ' ') --slen;
  *len = slen;
  *dest = NULL;
  if (slen > 0) {
    *dest = (char *)malloc(slen);
    memcpy(*dest, src, slen);
  }
}
void _gfortran_set_args(int argc, char **argv) {
  (void)argc;
  (void)argv;
}
void _gfortran_set_options(int n, int *opts) {
  (void)n;
  (void)opts;
}
static void flush_all(void) {
  for (int k = 0; k < MAXOPEN; ++k)
    if (units[k].used && units[k].fp) fflush(units[k].fp);
}
void _gfortran_stop_string(const char *s, int len) {
  if (s && len > 0) {
    fprintf(stderr, "STOP ");
    fwrite(s, 1, len, stderr);
    fprintf(stderr, "\n");
  }
  flush_all();
  exit(0);
}
void _gfortran_runtime_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Fortran runtime error: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  flush_all();
  exit(2);
}
void _gfortran_runtime_error_at(const char *where, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "%s\nFortran runtime error: ", where);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  flush_all();
  exit(2);
}
void _gfortran_os_error(const char *msg) {
  fprintf(stderr, "Operating system error: %s\n", msg);
  flush_all();
  exit(1);
}
