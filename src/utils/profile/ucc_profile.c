#include "ucc_profile.h"
#include "utils/ucc_log.h"
#include "utils/ucc_atomic.h"
#include <dlfcn.h>
#include <stdio.h>
#include <unistd.h>

unsigned ucc_profile_mode_mask = 0; /* bit0 accum, bit1 log */
typedef struct prof_rec { const ucc_profile_loc_t *loc; double t0, t1; const void *req; } prof_rec_t;
static prof_rec_t *recs; static size_t n_recs, cap_recs, rec_pos;
static ucc_profile_loc_t *locs[512]; static int n_locs;
static char out_file[512];
static ucc_spinlock_t lock;
static int  (*nvtx_push)(const char *); static int (*nvtx_pop)(void);

void ucc_profile_init(const char *mode_str, const char *file, size_t log_size)
{
    void *h;
    ucc_profile_mode_mask = 0;
    if (mode_str && strstr(mode_str, "accum")) ucc_profile_mode_mask |= 1;
    if (mode_str && strstr(mode_str, "log")) ucc_profile_mode_mask |= 2;
    if (!ucc_profile_mode_mask) return;
    out_file[0] = 0;
    for (const char *p = file ? file : "ucc_%h_%p.prof"; *p && strlen(out_file) + 64 < sizeof(out_file); p++) {
        size_t o = strlen(out_file);
        if (p[0] == '%' && p[1] == 'p') { snprintf(out_file + o, sizeof(out_file) - o, "%d", (int)getpid()); p++; }
        else if (p[0] == '%' && p[1] == 'h') { snprintf(out_file + o, sizeof(out_file) - o, "%s", ucc_get_host_name()); p++; }
        else { out_file[o] = *p; out_file[o + 1] = 0; }
    }
    if (ucc_profile_mode_mask & 2) { cap_recs = log_size / sizeof(prof_rec_t); if (cap_recs < 16) cap_recs = 16; recs = (prof_rec_t *)calloc(cap_recs, sizeof(prof_rec_t)); }
    h = dlopen("libnvToolsExt.so.1", RTLD_LAZY | RTLD_GLOBAL);
    if (h) { nvtx_push = (int (*)(const char *))dlsym(h, "nvtxRangePushA"); nvtx_pop = (int (*)(void))dlsym(h, "nvtxRangePop"); }
}

void ucc_profile_record(ucc_profile_loc_t *loc, double t0, double t1, const void *req)
{
    ucc_spin_lock(&lock);
    if (!loc->registered && n_locs < 512) { locs[n_locs++] = loc; loc->registered = 1; }
    loc->count++; loc->total += t1 - t0;
    if ((ucc_profile_mode_mask & 2) && recs) { /* ring: new records replace old ones */
        recs[rec_pos].loc = loc; recs[rec_pos].t0 = t0; recs[rec_pos].t1 = t1; recs[rec_pos].req = req;
        rec_pos = (rec_pos + 1) % cap_recs; if (n_recs < cap_recs) n_recs++;
    }
    ucc_spin_unlock(&lock);
}

typedef struct dyn_loc { ucc_profile_loc_t loc; char name[64]; } dyn_loc_t;
static dyn_loc_t *dyn_locs[128]; static int n_dyn;
void ucc_profile_event_named(const char *prefix, const char *mid, const char *suffix, const void *req)
{
    char name[64]; dyn_loc_t *d = NULL; double t;
    snprintf(name, sizeof(name), "%s_%s_%s", prefix, mid, suffix);
    for (char *c = name; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    ucc_spin_lock(&lock);
    for (int i = 0; i < n_dyn; i++) if (!strcmp(dyn_locs[i]->name, name)) { d = dyn_locs[i]; break; }
    if (!d && n_dyn < 128 && (d = (dyn_loc_t *)calloc(1, sizeof(*d)))) {
        memcpy(d->name, name, sizeof(name)); d->loc.name = d->name; d->loc.file = "-"; d->loc.line = 0; dyn_locs[n_dyn++] = d;
    }
    ucc_spin_unlock(&lock);
    if (!d) return;
    t = ucc_get_time();
    ucc_profile_record(&d->loc, t, t, req);
}

void ucc_profile_cleanup(void)
{
    FILE *f;
    if (!ucc_profile_mode_mask) return;
    f = fopen(out_file, "w");
    if (f) {
        fprintf(f, "# ucc_b200 profile pid %d host %s\n", (int)getpid(), ucc_get_host_name());
        if (ucc_profile_mode_mask & 1) {
            fprintf(f, "# accum: name file:line count total_us avg_us\n");
            for (int i = 0; i < n_locs; i++)
                fprintf(f, "A %s %s:%d %lu %.3f %.3f\n", locs[i]->name, locs[i]->file, locs[i]->line, (unsigned long)locs[i]->count,
                        locs[i]->total * 1e6, locs[i]->count ? locs[i]->total * 1e6 / (double)locs[i]->count : 0.0);
        }
        if ((ucc_profile_mode_mask & 2) && recs) {
            size_t start = n_recs < cap_recs ? 0 : rec_pos;
            fprintf(f, "# log: name t_begin_s dur_us req\n");
            for (size_t k = 0; k < n_recs; k++) { prof_rec_t *r = &recs[(start + k) % cap_recs];
                fprintf(f, "L %s %.9f %.3f %p\n", r->loc->name, r->t0, (r->t1 - r->t0) * 1e6, r->req); }
        }
        fclose(f);
    }
    free(recs); recs = NULL; n_recs = cap_recs = rec_pos = 0; ucc_profile_mode_mask = 0;
    for (int i = 0; i < n_dyn; i++) free(dyn_locs[i]);
    n_dyn = 0; n_locs = 0;
}
void ucc_profile_range_push(const char *name) { if (nvtx_push) nvtx_push(name); }
void ucc_profile_range_pop(void) { if (nvtx_pop) nvtx_pop(); }
