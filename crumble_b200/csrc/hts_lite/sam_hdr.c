/*
 * hts_lite: @PG header line insertion (see htslib/cram.h in this directory).
 * Behaviour follows the SAM specification for @PG chaining: the new record gets a
 * unique ID derived from the program name and PP pointing at the last @PG in the
 * existing chain.
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "htslib/cram.h"

SAM_hdr *sam_hdr_parse_(const char *hdr, int len) {
    SAM_hdr *sh = (SAM_hdr *)calloc(1, sizeof(*sh));
    if (!sh) return NULL;
    sh->cap = (size_t)len + 1024;
    sh->text = (char *)malloc(sh->cap);
    memcpy(sh->text, hdr, (size_t)len);
    sh->len = (size_t)len;
    sh->text[sh->len] = 0;
    return sh;
}

static int id_in_use(const SAM_hdr *sh, const char *id) {
    char pat[300];
    snprintf(pat, sizeof pat, "\tID:%s", id);
    const char *p = sh->text;
    size_t pl = strlen(pat);
    while ((p = strstr(p, "@PG")) != NULL) {
        const char *eol = strchr(p, '\n');
        if (!eol) eol = sh->text + sh->len;
        const char *q = p;
        while ((q = strstr(q, pat)) != NULL && q < eol) {
            if (q[pl] == '\t' || q[pl] == '\n' || q[pl] == 0) return 1;
            q += pl;
        }
        p = eol;
    }
    return 0;
}

/* ID of the last @PG line that no other @PG names in a PP tag */
static int chain_tail(const SAM_hdr *sh, char *out, size_t outsz) {
    const char *p = sh->text; int found = 0;
    while ((p = strstr(p, "@PG\t")) != NULL) {
        if (p != sh->text && p[-1] != '\n') { p += 3; continue; }
        const char *eol = strchr(p, '\n');
        if (!eol) eol = sh->text + sh->len;
        const char *id = strstr(p, "\tID:");
        if (id && id < eol) {
            id += 4;
            size_t l = strcspn(id, "\t\n");
            char tmp[256], pat[300];
            if (l < sizeof tmp) {
                memcpy(tmp, id, l); tmp[l] = 0;
                snprintf(pat, sizeof pat, "\tPP:%s", tmp);
                const char *u = strstr(sh->text, pat);
                int used = 0;
                while (u) { char e = u[strlen(pat)]; if (e == '\t' || e == '\n' || e == 0) { used = 1; break; } u = strstr(u + 1, pat); }
                if (!used) { snprintf(out, outsz, "%s", tmp); found = 1; }
            }
        }
        p = eol;
    }
    return found;
}

int sam_hdr_add_PG(SAM_hdr *sh, const char *name, ...) {
    char id[280], pp[256];
    snprintf(id, sizeof id, "%s", name);
    for (int n = 1; id_in_use(sh, id); n++) snprintf(id, sizeof id, "%s.%d", name, n);
    int have_pp = chain_tail(sh, pp, sizeof pp);

    size_t cap = 1024, n = 0;
    char *line = (char *)malloc(cap);
    n += (size_t)snprintf(line + n, cap - n, "@PG\tID:%s\tPN:%s", id, name);
    if (have_pp) n += (size_t)snprintf(line + n, cap - n, "\tPP:%s", pp);
    va_list ap; va_start(ap, name);
    for (;;) {
        const char *k = va_arg(ap, const char *);
        if (!k) break;
        const char *v = va_arg(ap, const char *);
        if (!v) break;
        size_t need = strlen(k) + strlen(v) + 8;
        if (n + need > cap) { cap = (n + need) * 2; line = (char *)realloc(line, cap); }
        n += (size_t)snprintf(line + n, cap - n, "\t%s:%s", k, v);
    }
    va_end(ap);
    if (sh->len && sh->text[sh->len - 1] != '\n') { sh->text[sh->len++] = '\n'; }
    if (sh->len + n + 2 > sh->cap) { sh->cap = (sh->len + n + 2) * 2; sh->text = (char *)realloc(sh->text, sh->cap); }
    memcpy(sh->text + sh->len, line, n); sh->len += n;
    sh->text[sh->len++] = '\n'; sh->text[sh->len] = 0;
    free(line);
    return 0;
}

const char *sam_hdr_str(SAM_hdr *sh) { return sh->text; }
int sam_hdr_length(SAM_hdr *sh) { return (int)sh->len; }
void sam_hdr_free(SAM_hdr *sh) { if (sh) { free(sh->text); free(sh); } }

char *stringify_argv(int argc, char *argv[]) {
    size_t n = 1;
    for (int i = 0; i < argc; i++) n += strlen(argv[i]) + 1;
    char *s = (char *)malloc(n), *w = s;
    for (int i = 0; i < argc; i++) {
        if (i) *w++ = ' ';
        for (const char *p = argv[i]; *p; p++) *w++ = (*p == '\t') ? ' ' : *p;
    }
    *w = 0;
    return s;
}
