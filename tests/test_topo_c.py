"""Topology / sub-group (sbgp) unit tests from a compiled C program on synthetic process tables (model: reference
test/gtest/core/test_topo.cc): node / socket / numa groups, leaders, rails (NET), host-ordered full group, subsets,
node-leader table, NVLink predicates from a fake GPU table."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "core/ucc_context.h"
#include "components/topo/ucc_topo.h"
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); exit(1); } } while (0)

/* storage of n fake context addresses: host h[i], socket s[i] (numa = socket), optional GPU index */
static ucc_addr_storage_t *mk_storage(int n, const int *host, const int *sock, int with_gpus, int nvswitch)
{
    ucc_addr_storage_t *st = calloc(1, sizeof(*st));
    st->addr_len = sizeof(ucc_context_addr_header_t); st->size = n; st->rank = 0;
    st->storage = calloc(n, st->addr_len);
    for (int i = 0; i < n; i++) {
        ucc_context_addr_header_t *h = UCC_ADDR_STORAGE_RANK_HEADER(st, i);
        h->ctx_id.pi.host_hash = 1000 + host[i]; h->ctx_id.pi.pid = 100 + i;
        h->ctx_id.pi.socket_id = sock ? sock[i] : UCC_SOCKET_ID_INVALID; h->ctx_id.pi.numa_id = sock ? sock[i] : UCC_NUMA_ID_INVALID;
        h->host_info.current_gpu = -1;
        if (with_gpus) {
            int local = 0; for (int q = 0; q < i; q++) if (host[q] == host[i]) local++;
            h->host_info.n_gpus = 4; h->host_info.current_gpu = local % 4;
            for (int g = 0; g < 4; g++) { h->host_info.gpus[g].uuid_hash = 7000 + host[i] * 10 + g; if (nvswitch) h->host_info.gpus[g].caps |= UCC_GPU_CAP_NVSWITCH; }
            /* without a switch: GPUs 0-1 and 2-3 are NVLink pairs */
            if (!nvswitch) { h->host_info.nvlink_matrix[0][1] = h->host_info.nvlink_matrix[1][0] = 2; h->host_info.nvlink_matrix[2][3] = h->host_info.nvlink_matrix[3][2] = 2; }
        }
    }
    return st;
}
static int has(const ucc_sbgp_t *s, ucc_rank_t r) { for (ucc_rank_t i = 0; i < s->group_size; i++) if (s->rank_map[i] == r) return 1; return 0; }
static void expect(const ucc_sbgp_t *s, int n, const int *ranks, int line)
{
    if (s->status != UCC_SBGP_ENABLED || (int)s->group_size != n) { printf("FAIL line %d: sbgp %s status %d size %u (expected %d)\n", line, ucc_sbgp_str(s->type), s->status, s->group_size, n); exit(1); }
    for (int i = 0; i < n; i++) if ((int)s->rank_map[i] != ranks[i]) { printf("FAIL line %d: sbgp %s member %d is %u, expected %d\n", line, ucc_sbgp_str(s->type), i, s->rank_map[i], ranks[i]); exit(1); }
    for (int i = 0; i < n; i++) if ((int)ucc_ep_map_eval(s->map, i) != ranks[i]) { printf("FAIL line %d: map mismatch\n", line); exit(1); }
}
#define EXPECT(_s, ...) do { int _r[] = {__VA_ARGS__}; expect(_s, (int)(sizeof(_r) / sizeof(_r[0])), _r, __LINE__); } while (0)

static ucc_topo_t *team_topo(ucc_context_topo_t *ct, int n, int me) { ucc_subset_t s; ucc_topo_t *t; s.map = ucc_ep_map_create_full(n); s.myrank = me; CHECK(ucc_topo_init(s, ct, &t) == UCC_OK); return t; }

/* 3 nodes, ppn 4 / 4 / 2, two sockets per node */
static void test_blocked(void)
{
    int host[10] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2}, sock[10] = {0, 0, 1, 1, 0, 0, 1, 1, 0, 1};
    ucc_addr_storage_t *st = mk_storage(10, host, sock, 0, 0);
    ucc_context_topo_t *ct; ucc_topo_t *t; ucc_sbgp_t *all; int n; ucc_rank_t *nl;
    CHECK(ucc_context_topo_init(st, &ct) == UCC_OK);
    CHECK(ct->nnodes == 3 && ct->min_ppn == 2 && ct->max_ppn == 4 && ct->sock_bound && ct->numa_bound);
    t = team_topo(ct, 10, 5);
    CHECK(ucc_topo_nnodes(t) == 3 && ucc_topo_min_ppn(t) == 2 && ucc_topo_max_ppn(t) == 4 && !ucc_topo_isoppn(t) && !ucc_topo_is_single_node(t));
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE), 4, 5, 6, 7);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_NODE)->group_rank == 1);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_SOCKET), 4, 5);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NUMA), 4, 5);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_LEADERS)->status == UCC_SBGP_NOT_EXISTS);   /* rank 5 is no leader */
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_SOCKET_LEADERS)->status == UCC_SBGP_NOT_EXISTS);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_NET)->status == UCC_SBGP_NOT_EXISTS);            /* unequal ppn: no rails */
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_FULL), 0, 1, 2, 3, 4, 5, 6, 7, 8, 9);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_FULL)->group_rank == 5);
    CHECK(ucc_topo_get_all_nodes(t, &all, &n) == UCC_OK && n == 3 && all[0].group_size == 4 && all[2].group_size == 2 && all[1].group_rank == 1 && all[0].group_rank == UCC_RANK_INVALID);
    CHECK(ucc_topo_get_all_sockets(t, &all, &n) == UCC_OK && n == 2 && has(&all[0], 4) && has(&all[1], 7));
    CHECK(ucc_topo_get_node_leaders(t, &nl) == UCC_OK && nl[0] == 0 && nl[3] == 0 && nl[5] == 4 && nl[9] == 8 && t->node_leader_rank == 4);
    CHECK(t->min_socket_size == 1 && t->max_socket_size == 2);
    CHECK(ucc_topo_min_socket_size(t) == 1 && ucc_topo_max_socket_size(t) == 2 && ucc_topo_min_numa_size(t) == 1 && ucc_topo_max_numa_size(t) == 2);
    CHECK(!ucc_topo_is_single_ppn(t) && ucc_topo_n_numas(t) == 2 && ucc_topo_n_sockets(t) == 2);
    CHECK(ucc_topo_get_node_host_id(t, 0) == 0 && ucc_topo_get_node_host_id(t, 6) == 1 && ucc_topo_get_node_host_id(t, 9) == 2);
    { ucc_subset_t ss = ucc_sbgp_to_subset(ucc_topo_get_sbgp(t, UCC_SBGP_NODE)); CHECK(ss.myrank == 1 && ss.map.ep_num == 4 && ucc_ep_map_eval(ss.map, 3) == 7); }
    CHECK(ucc_topo_get_all_node_nvlinks(t, &all, &n) == UCC_ERR_NOT_FOUND);   /* no device information */
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_LAST) == NULL);
    ucc_topo_cleanup(t);
    t = team_topo(ct, 10, 4);   /* a node leader and socket leader */
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_LEADERS), 0, 4, 8);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_LEADERS)->group_rank == 1);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_SOCKET_LEADERS), 4, 6);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NUMA_LEADERS), 4, 6);
    CHECK(!ucc_topo_has_device_info(t));
    ucc_topo_cleanup(t);
    ucc_context_topo_cleanup(ct); free(st->storage); free(st);
}

/* 3 nodes x 3, ranks dealt round-robin over the hosts; no socket binding */
static void test_round_robin(void)
{
    int host[9] = {0, 1, 2, 0, 1, 2, 0, 1, 2};
    ucc_addr_storage_t *st = mk_storage(9, host, NULL, 0, 0);
    ucc_context_topo_t *ct; ucc_topo_t *t; ucc_subset_t sub; ucc_rank_t evens[5] = {0, 2, 4, 6, 8}, *arr = evens;
    CHECK(ucc_context_topo_init(st, &ct) == UCC_OK);
    CHECK(ct->nnodes == 3 && ct->min_ppn == 3 && ct->max_ppn == 3 && !ct->sock_bound);
    t = team_topo(ct, 9, 4);
    CHECK(ucc_topo_isoppn(t) && ucc_topo_n_numas(t) == 0 && ucc_topo_get_node_host_id(t, 5) == 2);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE), 1, 4, 7);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NET), 3, 4, 5);                       /* second process of every node */
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_SOCKET)->status == UCC_SBGP_NOT_EXISTS);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_FULL_HOST_ORDERED), 0, 3, 6, 1, 4, 7, 2, 5, 8);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_FULL_HOST_ORDERED)->group_rank == 4);
    ucc_topo_cleanup(t);
    /* sub-team of the even context ranks: team rank i = ctx rank 2i -> hosts 0 2 1 0 2 */
    sub.map = ucc_ep_map_from_array(&arr, 5, 9, 0); sub.myrank = 3;
    CHECK(ucc_topo_init(sub, ct, &t) == UCC_OK);
    CHECK(ucc_topo_nnodes(t) == 3 && ucc_topo_min_ppn(t) == 1 && ucc_topo_max_ppn(t) == 2);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE), 0, 3);
    CHECK(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_LEADERS)->status == UCC_SBGP_NOT_EXISTS);
    ucc_topo_cleanup(t);
    sub.myrank = 1;
    CHECK(ucc_topo_init(sub, ct, &t) == UCC_OK);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_LEADERS), 0, 1, 2);
    EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE), 1, 4);
    ucc_topo_cleanup(t);
    ucc_context_topo_cleanup(ct); free(st->storage); free(st);
}

static void test_nvlink(void)
{
    int host[8] = {0, 0, 0, 0, 1, 1, 1, 1}, sock[8] = {0, 0, 1, 1, 0, 0, 1, 1};
    for (int nvswitch = 0; nvswitch < 2; nvswitch++) {
        ucc_addr_storage_t *st = mk_storage(8, host, sock, 1, nvswitch);
        ucc_context_topo_t *ct; ucc_topo_t *t;
        CHECK(ucc_context_topo_init(st, &ct) == UCC_OK);
        t = team_topo(ct, 8, 2);
        CHECK(ucc_topo_has_device_info(t) && ucc_topo_rank_gpu(t, 2, NULL) == 2 && ucc_topo_rank_gpu(t, 5, NULL) == 1);
        CHECK(!ucc_topo_is_nvlink_fully_connected(t) && !ucc_topo_is_single_nvlink_domain(t));   /* two nodes, no fabric */
        CHECK(ucc_topo_nvlink_connected(t, 2, 3) && !ucc_topo_nvlink_connected(t, 2, 6));
        if (nvswitch) { CHECK(ucc_topo_nvlink_connected(t, 0, 3)); EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_NVLINK), 0, 1, 2, 3); }
        else { CHECK(!ucc_topo_nvlink_connected(t, 0, 3)); EXPECT(ucc_topo_get_sbgp(t, UCC_SBGP_NODE_NVLINK), 2, 3); }
        { ucc_sbgp_t *isl; int ni;   /* NVLink islands of my node: one with a switch, the two pairs without */
          CHECK(ucc_topo_get_all_node_nvlinks(t, &isl, &ni) == UCC_OK && ni == (nvswitch ? 1 : 2));
          if (nvswitch) { EXPECT(&isl[0], 0, 1, 2, 3); CHECK(isl[0].group_rank == 2); }
          else { EXPECT(&isl[0], 0, 1); EXPECT(&isl[1], 2, 3); CHECK(isl[0].group_rank == UCC_RANK_INVALID && isl[1].group_rank == 0); } }
        ucc_topo_cleanup(t);
        /* one node only: with a switch the 4 GPUs are one NVLink domain */
        { ucc_subset_t sub; ucc_rank_t first[4] = {0, 1, 2, 3}, *arr = first; sub.map = ucc_ep_map_from_array(&arr, 4, 8, 0); sub.myrank = 0;
          CHECK(ucc_topo_init(sub, ct, &t) == UCC_OK);
          CHECK(ucc_topo_is_single_node(t) && ucc_topo_is_nvlink_fully_connected(t) == nvswitch && ucc_topo_is_single_nvlink_domain(t) == nvswitch);
          ucc_topo_cleanup(t); }
        ucc_context_topo_cleanup(ct); free(st->storage); free(st);
    }
}

int main(void) { test_blocked(); test_round_robin(); test_nvlink(); printf("TOPO_OK\n"); return 0; }
'''


def test_topo_compiled(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    libdir = os.path.dirname(os.environ.get("UCC_B200_LIB") or os.path.join(ROOT, "ucc_b200", "lib", "libucc.so"))
    cc = subprocess.run(["gcc", "-O1", "-g", "-std=gnu11", "-Wall", "-Wno-unused-result", f"-I{ROOT}/src", f"-I{ROOT}/include", "-D_GNU_SOURCE", str(src), "-o", str(exe),
                         f"-L{libdir}", "-lucc", f"-Wl,-rpath,{libdir}", "-lpthread"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-4000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert "TOPO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
