"""Internal C building blocks unit-tested from a compiled C program linked against libucc.so (model: reference
test/gtest/utils/*, test/gtest/core/test_schedule.cc, test_mpool etc.): mpool, lock-free queue, ep maps, block math,
task/event manager, schedules with dependencies and error propagation, pipelined schedules in the three orders."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "utils/ucc_mpool.h"
#include "utils/ucc_lock_free_queue.h"
#include "utils/ucc_coll_utils.h"
#include "utils/ucc_math.h"
#include "utils/ucc_string.h"
#include "schedule/ucc_schedule.h"
#include "schedule/ucc_schedule_pipelined.h"
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); exit(1); } } while (0)

/* ---------------- dummy tasks: post either completes inline or parks the task until the driver completes it */
typedef struct dtask { ucc_coll_task_t super; int id, frag, defer, posts; ucc_status_t result; } dtask_t;
static dtask_t *parked[256]; static int n_parked;
static char     logbuf[8192];
static int      start_order[64][16], done_order[64][16], clock_;
static ucc_status_t d_post(ucc_coll_task_t *t)
{
    dtask_t *d = (dtask_t *)t;
    d->posts++;
    t->status = UCC_INPROGRESS; t->super.status = UCC_INPROGRESS;
    start_order[d->frag][d->id] = ++clock_;
    if (d->defer) { parked[n_parked++] = d; return UCC_OK; }
    done_order[d->frag][d->id] = ++clock_;
    t->status = d->result;
    return ucc_task_complete(t) < 0 ? UCC_OK : UCC_OK;
}
static ucc_status_t d_fin(ucc_coll_task_t *t) { ucc_coll_task_destruct(t); free(t); return UCC_OK; }
static dtask_t *mk(int id, int defer) { dtask_t *d = calloc(1, sizeof(*d)); ucc_coll_task_init(&d->super, NULL, NULL); d->super.post = d_post; d->super.finalize = d_fin; d->id = id; d->defer = defer; d->result = UCC_OK; return d; }
static void drain(void) { while (n_parked) { dtask_t *d = parked[0]; memmove(parked, parked + 1, sizeof(parked[0]) * (size_t)(--n_parked)); done_order[d->frag][d->id] = ++clock_; d->super.status = d->result; ucc_task_complete(&d->super); } }
static void drain_lifo(void) { while (n_parked) { dtask_t *d = parked[--n_parked]; done_order[d->frag][d->id] = ++clock_; d->super.status = d->result; ucc_task_complete(&d->super); } }

static void test_schedule(void)
{
    for (int defer = 0; defer < 2; defer++) {
        ucc_schedule_t s; dtask_t *t[4];
        ucc_schedule_init(&s, NULL, NULL);
        for (int i = 0; i < 4; i++) { t[i] = mk(i, defer); CHECK(ucc_schedule_add_task(&s, &t[i]->super) == UCC_OK); }
        /* diamond: 0 -> {1,2} -> 3 */
        ucc_task_subscribe_dep(&s.super, &t[0]->super, UCC_EVENT_SCHEDULE_STARTED);
        ucc_task_subscribe_dep(&t[0]->super, &t[1]->super, UCC_EVENT_COMPLETED); ucc_task_subscribe_dep(&t[0]->super, &t[2]->super, UCC_EVENT_COMPLETED);
        ucc_task_subscribe_dep(&t[1]->super, &t[3]->super, UCC_EVENT_COMPLETED); ucc_task_subscribe_dep(&t[2]->super, &t[3]->super, UCC_EVENT_COMPLETED);
        for (int rep = 0; rep < 3; rep++) { /* persistent style re-post */
            clock_ = 0; memset(start_order, 0, sizeof(start_order)); memset(done_order, 0, sizeof(done_order));
            CHECK(s.super.post(&s.super) == UCC_OK);
            if (defer) { CHECK(s.super.super.status == UCC_INPROGRESS); drain(); }
            CHECK(s.super.super.status == UCC_OK);
            CHECK(start_order[0][1] > done_order[0][0] && start_order[0][2] > done_order[0][0]);
            CHECK(start_order[0][3] > done_order[0][1] && start_order[0][3] > done_order[0][2]);
            for (int i = 0; i < 4; i++) CHECK(t[i]->posts == rep + 1);
        }
        s.super.finalize(&s.super);
    }
    { /* error in the middle propagates to the schedule; the dependent task never starts */
        ucc_schedule_t s; dtask_t *a = mk(0, 0), *b = mk(1, 0), *c = mk(2, 0);
        ucc_schedule_init(&s, NULL, NULL);
        ucc_schedule_add_task(&s, &a->super); ucc_schedule_add_task(&s, &b->super); ucc_schedule_add_task(&s, &c->super);
        ucc_task_subscribe_dep(&s.super, &a->super, UCC_EVENT_SCHEDULE_STARTED);
        ucc_task_subscribe_dep(&a->super, &b->super, UCC_EVENT_COMPLETED); ucc_task_subscribe_dep(&b->super, &c->super, UCC_EVENT_COMPLETED);
        b->result = UCC_ERR_NO_MESSAGE;
        freopen("/dev/null", "w", stderr);
        s.super.post(&s.super);
        CHECK(s.super.super.status == UCC_ERR_NO_MESSAGE); CHECK(c->posts == 0);
        s.super.finalize(&s.super);
    }
}

/* ---------------- pipelined schedule: fragments of 3 chained tasks */
static int n_frag_tasks = 3, frag_defer;
static ucc_status_t frag_init(ucc_base_coll_args_t *a, ucc_schedule_pipelined_t *sp, ucc_base_team_t *team, ucc_schedule_t **out)
{
    ucc_schedule_t *f = calloc(1, sizeof(*f)); (void)a; (void)sp; (void)team;
    ucc_schedule_init(f, NULL, NULL);
    f->super.finalize = ucc_schedule_finalize;
    for (int i = 0; i < n_frag_tasks; i++) {
        dtask_t *t = mk(i, frag_defer);
        ucc_schedule_add_task(f, &t->super);
        if (i == 0) ucc_task_subscribe_dep(&f->super, &t->super, UCC_EVENT_SCHEDULE_STARTED);
        else ucc_task_subscribe_dep(f->tasks[i - 1], &t->super, UCC_EVENT_COMPLETED);
    }
    *out = f;
    return UCC_OK;
}
static ucc_status_t frag_setup(ucc_schedule_pipelined_t *sp, ucc_schedule_t *f, int frag_num)
{ (void)sp; for (unsigned i = 0; i < f->n_tasks; i++) ((dtask_t *)f->tasks[i])->frag = frag_num; return UCC_OK; }

static void test_pipelined(void)
{
    for (int order = 0; order < 3; order++) for (int depth = 1; depth <= 4; depth++) for (int total = 1; total <= 9; total += 2) for (int mode = 0; mode < 3; mode++) {
        ucc_schedule_pipelined_t *sp = calloc(1, sizeof(*sp));
        frag_defer = mode > 0;
        CHECK(ucc_schedule_pipelined_init(NULL, NULL, frag_init, frag_setup, depth, total, (ucc_pipeline_order_t)order, sp) == UCC_OK);
        for (int rep = 0; rep < 2; rep++) {
            clock_ = 0; memset(start_order, 0, sizeof(start_order)); memset(done_order, 0, sizeof(done_order));
            CHECK(ucc_schedule_pipelined_post(&sp->super.super) >= 0);
            for (int guard = 0; guard < 1000 && sp->super.super.super.status == UCC_INPROGRESS; guard++) { if (mode == 2) drain_lifo(); else drain(); }
            CHECK(sp->super.super.super.status == UCC_OK);
            for (int g = 0; g < total; g++) for (int i = 0; i < n_frag_tasks; i++) {
                CHECK(start_order[g][i] > 0 && done_order[g][i] >= start_order[g][i]);
                if (i) CHECK(start_order[g][i] > done_order[g][i - 1]);                                     /* chain inside a fragment */
                if (g && order == UCC_PIPELINE_SEQUENTIAL && depth > 1) CHECK(start_order[g][i] > done_order[g - 1][i]);
                if (g && order == UCC_PIPELINE_ORDERED && depth > 1) CHECK(start_order[g][i] > start_order[g - 1][i]);
                if (g >= depth && i == 0) { /* never more than `depth` fragments in flight: g-depth+1 of them finished before g started */
                    int fin = 0; for (int h = 0; h < total; h++) if (done_order[h][n_frag_tasks - 1] && done_order[h][n_frag_tasks - 1] < start_order[g][0]) fin++;
                    CHECK(fin >= g - depth + 1); }
            }
        }
        ucc_schedule_pipelined_finalize(&sp->super.super);
        free(sp);
    }
}

/* ---------------- utils */
static void *lfq_producer(void *arg) { ucc_lf_queue_t *q = arg; for (int i = 0; i < 20000; i++) { ucc_lf_queue_elem_t *e = malloc(sizeof(*e) + sizeof(int)); ucc_lf_queue_enqueue(q, e); } return NULL; }
static void test_utils(void)
{
    ucc_mpool_t mp; void *objs[100];
    CHECK(ucc_mpool_init(&mp, 0, 48, 0, 16, 8, 64, NULL, UCC_THREAD_SINGLE, "t") == UCC_OK);
    for (int i = 0; i < 64; i++) { objs[i] = ucc_mpool_get(&mp); CHECK(objs[i] && ((uintptr_t)objs[i] % 16) == 0); memset(objs[i], i, 48); }
    CHECK(ucc_mpool_get(&mp) == NULL);                  /* max_elems reached */
    for (int i = 0; i < 64; i++) { CHECK(((unsigned char *)objs[i])[47] == i); ucc_mpool_put(objs[i]); }
    CHECK(ucc_mpool_get(&mp) != NULL);
    ucc_mpool_cleanup(&mp, 0);
    { ucc_lf_queue_t q; pthread_t th[4]; int got = 0; ucc_lf_queue_init(&q);
      for (int i = 0; i < 4; i++) pthread_create(&th[i], NULL, lfq_producer, &q);
      while (got < 80000) { ucc_lf_queue_elem_t *e = ucc_lf_queue_dequeue(&q); if (e) { free(e); got++; } }
      for (int i = 0; i < 4; i++) pthread_join(th[i], NULL);
      CHECK(ucc_lf_queue_dequeue(&q) == NULL); ucc_lf_queue_destroy(&q); }
    { ucc_ep_map_t full = ucc_ep_map_create_full(7), rev = ucc_ep_map_create_reverse(7), inv, strided; uint64_t arr[4] = {9, 3, 5, 1}; ucc_ep_map_t am;
      for (ucc_rank_t r = 0; r < 7; r++) { CHECK(ucc_ep_map_eval(full, r) == r); CHECK(ucc_ep_map_eval(rev, r) == 6 - r); CHECK(ucc_ep_map_local_rank(rev, 6 - r) == r); }
      memset(&strided, 0, sizeof(strided)); strided.type = UCC_EP_MAP_STRIDED; strided.ep_num = 4; strided.strided.start = 10; strided.strided.stride = -3;
      for (ucc_rank_t r = 0; r < 4; r++) { CHECK(ucc_ep_map_eval(strided, r) == 10 - 3 * r); CHECK(ucc_ep_map_local_rank(strided, 10 - 3 * r) == r); }
      CHECK(ucc_ep_map_local_rank(strided, 9) == UCC_RANK_INVALID);
      memset(&am, 0, sizeof(am)); am.type = UCC_EP_MAP_ARRAY; am.ep_num = 4; am.array.map = arr; am.array.elem_size = 8;
      for (ucc_rank_t r = 0; r < 4; r++) { CHECK(ucc_ep_map_eval(am, r) == arr[r]); CHECK(ucc_ep_map_local_rank(am, (ucc_rank_t)arr[r]) == r); }
      CHECK(ucc_ep_map_create_inverse(rev, &inv, 1) == UCC_OK); for (ucc_rank_t r = 0; r < 7; r++) CHECK(ucc_ep_map_eval(inv, ucc_ep_map_eval(rev, r)) == r); }
    for (size_t total = 0; total < 50; total++) for (unsigned n = 1; n < 9; n++) { size_t sum = 0; for (unsigned i = 0; i < n; i++) { CHECK(ucc_buffer_block_offset(total, n, i) == sum); sum += ucc_buffer_block_count(total, n, i); } CHECK(sum == total); }
    { size_t v; CHECK(ucc_str_to_memunits("3m", &v) == UCC_OK && v == 3u << 20); CHECK(ucc_str_to_memunits("x", &v) != UCC_OK); }
}

int main(void) { test_utils(); test_pipelined(); test_schedule(); printf("INTERNALS_OK\n"); return 0; }
'''


def test_internals_compiled(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "ucc_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libucc.so")):
        subprocess.run(["make", "-C", ROOT, "-j8", "core"], capture_output=True)
    cc = subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-Wno-unused-result", f"-I{ROOT}/src", f"-I{ROOT}/include", "-D_GNU_SOURCE", str(src), "-o", str(exe),
                         f"-L{libdir}", "-lucc", f"-Wl,-rpath,{libdir}", "-lpthread"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-4000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert "INTERNALS_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
