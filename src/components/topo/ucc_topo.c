#include "ucc_topo.h"
#include "core/ucc_context.h"
#include "utils/ucc_log.h"
#include "utils/ucc_math.h"

/* ---------------- context topo ---------------- */
ucc_status_t ucc_context_topo_init(ucc_addr_storage_t *storage, ucc_context_topo_t **topo_p)
{
    ucc_context_topo_t *t = (ucc_context_topo_t *)calloc(1, sizeof(*t));
    ucc_rank_t n = storage->size, nhosts = 0;
    ucc_host_id_t *hh;
    ucc_rank_t *ppn;
    if (!t) return UCC_ERR_NO_MEMORY;
    t->n_procs = n;
    t->procs = (ucc_proc_info_t *)calloc(n, sizeof(ucc_proc_info_t));
    t->hosts = (ucc_host_info_t *)calloc(n, sizeof(ucc_host_info_t));
    hh = (ucc_host_id_t *)calloc(n, sizeof(*hh)); ppn = (ucc_rank_t *)calloc(n, sizeof(*ppn));
    if (!t->procs || !t->hosts || !hh || !ppn) { free(hh); free(ppn); ucc_context_topo_cleanup(t); return UCC_ERR_NO_MEMORY; }
    t->sock_bound = t->numa_bound = 1;
    for (ucc_rank_t i = 0; i < n; i++) {
        ucc_context_addr_header_t *h = UCC_ADDR_STORAGE_RANK_HEADER(storage, i);
        ucc_rank_t k;
        t->procs[i] = h->ctx_id.pi; t->hosts[i] = h->host_info;
        for (k = 0; k < nhosts; k++) if (hh[k] == t->procs[i].host_hash) break;
        if (k == nhosts) hh[nhosts++] = t->procs[i].host_hash;
        t->procs[i].host_id = k; ppn[k]++;
        if (t->procs[i].socket_id == UCC_SOCKET_ID_INVALID) t->sock_bound = 0;
        if (t->procs[i].numa_id == UCC_NUMA_ID_INVALID) t->numa_bound = 0;
    }
    t->nnodes = nhosts; t->min_ppn = UCC_RANK_MAX; t->max_ppn = 0;
    for (ucc_rank_t k = 0; k < nhosts; k++) { if (ppn[k] < t->min_ppn) t->min_ppn = ppn[k]; if (ppn[k] > t->max_ppn) t->max_ppn = ppn[k]; }
    free(hh); free(ppn);
    *topo_p = t;
    return UCC_OK;
}
void ucc_context_topo_cleanup(ucc_context_topo_t *t) { if (!t) return; free(t->procs); free(t->hosts); free(t); }

/* ---------------- team topo ---------------- */
static const char *sbgp_names[] = {"numa", "socket", "node", "node_nvlink", "node_leaders", "net", "socket_leaders", "numa_leaders", "full", "full_host_ordered"};
const char *ucc_sbgp_str(ucc_sbgp_type_t type) { return type < UCC_SBGP_LAST ? sbgp_names[type] : "unknown"; }

#define PROC(_topo, _r) (&(_topo)->topo->procs[ucc_ep_map_eval((_topo)->set.map, (_r))])
#define HOST(_topo, _r) (&(_topo)->topo->hosts[ucc_ep_map_eval((_topo)->set.map, (_r))])

static void sbgp_finish(ucc_sbgp_t *s, ucc_topo_t *topo, ucc_rank_t *ranks, ucc_rank_t n, int contains_me)
{
    s->rank_map = ranks; s->group_size = n; s->group_rank = UCC_RANK_INVALID;
    if (n == 0 || !contains_me) { free(ranks); s->rank_map = NULL; s->status = UCC_SBGP_NOT_EXISTS; return; }
    for (ucc_rank_t i = 0; i < n; i++) if (ranks[i] == topo->set.myrank) s->group_rank = i;
    /* map construction keeps rank_map alive: need_free = 0 */
    s->map = ucc_ep_map_from_array(&s->rank_map, n, (ucc_rank_t)topo->set.map.ep_num, 0);
    s->status = UCC_SBGP_ENABLED;
}

typedef int (*same_fn_t)(const ucc_topo_t *, ucc_rank_t, ucc_rank_t);
static int same_node(const ucc_topo_t *t, ucc_rank_t a, ucc_rank_t b) { return PROC(t, a)->host_hash == PROC(t, b)->host_hash; }
static int same_socket(const ucc_topo_t *t, ucc_rank_t a, ucc_rank_t b)
{ return same_node(t, a, b) && PROC(t, a)->socket_id == PROC(t, b)->socket_id && PROC(t, a)->socket_id != UCC_SOCKET_ID_INVALID; }
static int same_numa(const ucc_topo_t *t, ucc_rank_t a, ucc_rank_t b)
{ return same_node(t, a, b) && PROC(t, a)->numa_id == PROC(t, b)->numa_id && PROC(t, a)->numa_id != UCC_NUMA_ID_INVALID; }

/* members of my group under equivalence `same` */
static void build_local_group(ucc_topo_t *topo, ucc_sbgp_t *s, same_fn_t same)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, n = 0, me = topo->set.myrank;
    ucc_rank_t *ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
    for (ucc_rank_t r = 0; r < size; r++) if (r == me || same(topo, r, me)) ranks[n++] = r;
    sbgp_finish(s, topo, ranks, n, 1);
    if (s->status == UCC_SBGP_ENABLED && n == 1 && s->type != UCC_SBGP_NODE) s->status = UCC_SBGP_ENABLED; /* singleton groups are legal */
}
/* first (leader_id-th) member of every group; `of_node` restricts to my node */
static void build_leaders(ucc_topo_t *topo, ucc_sbgp_t *s, same_fn_t same, int only_my_node, ucc_rank_t leader_id)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, n = 0, me = topo->set.myrank;
    ucc_rank_t *ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
    int contains_me = 0;
    for (ucc_rank_t r = 0; r < size; r++) {
        ucc_rank_t before = 0, group_size = 0, pick;
        if (only_my_node && !same_node(topo, r, me)) continue;
        for (ucc_rank_t q = 0; q < size; q++) if (q == r || same(topo, q, r)) { if (q < r) before++; group_size++; }
        pick = leader_id < group_size ? leader_id : group_size - 1; /* small groups: last member leads */
        if (before == pick) { ranks[n++] = r; if (r == me) contains_me = 1; }
    }
    sbgp_finish(s, topo, ranks, n, contains_me);
}

static void build_full_host_ordered(ucc_topo_t *topo, ucc_sbgp_t *s)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, n = 0;
    ucc_rank_t *ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
    uint8_t *done = (uint8_t *)calloc(size, 1);
    /* stable grouping: hosts in order of first appearance, then socket, then rank */
    for (ucc_rank_t r = 0; r < size; r++) {
        if (done[r]) continue;
        for (int sock = -1; sock < 256; sock++) {
            for (ucc_rank_t q = r; q < size; q++) {
                if (done[q] || !same_node(topo, q, r)) continue;
                int qs = PROC(topo, q)->socket_id == UCC_SOCKET_ID_INVALID ? -1 : (int)PROC(topo, q)->socket_id;
                if (qs != sock) continue;
                ranks[n++] = q; done[q] = 1;
            }
        }
    }
    free(done);
    sbgp_finish(s, topo, ranks, n, 1);
}

int ucc_topo_rank_gpu(const ucc_topo_t *topo, ucc_rank_t r, ucc_gpu_info_t **gpu)
{
    ucc_host_info_t *h = HOST(topo, r);
    int idx = h->current_gpu;
    if (idx < 0 && h->visible_gpus && ucc_is_pow2(h->visible_gpus)) idx = (int)ucc_ilog2(h->visible_gpus); /* one visible GPU per process */
    if (idx < 0 || idx >= h->n_gpus) return -1;
    if (gpu) *gpu = &h->gpus[idx];
    return idx;
}
int ucc_topo_nvlink_connected(const ucc_topo_t *topo, ucc_rank_t a, ucc_rank_t b)
{
    ucc_gpu_info_t *ga, *gb;
    int ia = ucc_topo_rank_gpu(topo, a, &ga), ib = ucc_topo_rank_gpu(topo, b, &gb);
    if (ia < 0 || ib < 0) return 0;
    if (same_node(topo, a, b)) {
        if (ga->uuid_hash == gb->uuid_hash && ia == ib) return 1; /* same device */
        if ((ga->caps & UCC_GPU_CAP_NVSWITCH) && (gb->caps & UCC_GPU_CAP_NVSWITCH)) return 1;
        return HOST(topo, a)->nvlink_matrix[ia][ib] > 0;
    }
    return (ga->caps & UCC_GPU_CAP_FABRIC) && (gb->caps & UCC_GPU_CAP_FABRIC) && ga->fabric_partition == gb->fabric_partition &&
           ga->fabric_clique_id == gb->fabric_clique_id && ga->fabric_partition != 0;
}
int ucc_topo_has_device_info(const ucc_topo_t *topo)
{ for (ucc_rank_t r = 0; r < (ucc_rank_t)topo->set.map.ep_num; r++) if (ucc_topo_rank_gpu(topo, r, NULL) < 0) return 0; return 1; }
int ucc_topo_is_nvlink_fully_connected(const ucc_topo_t *topo)
{
    ucc_rank_t n = (ucc_rank_t)topo->set.map.ep_num;
    for (ucc_rank_t a = 0; a < n; a++) for (ucc_rank_t b = a + 1; b < n; b++) if (!ucc_topo_nvlink_connected(topo, a, b)) return 0;
    return 1;
}
int ucc_topo_is_single_nvlink_domain(const ucc_topo_t *topo)
{ return ucc_topo_has_device_info(topo) && ucc_topo_is_nvlink_fully_connected(topo); }

static void build_node_nvlink(ucc_topo_t *topo, ucc_sbgp_t *s)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, n = 0, me = topo->set.myrank;
    ucc_rank_t *ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
    for (ucc_rank_t r = 0; r < size; r++) if (r == me || (same_node(topo, r, me) && ucc_topo_nvlink_connected(topo, r, me))) ranks[n++] = r;
    sbgp_finish(s, topo, ranks, n, ucc_topo_rank_gpu(topo, me, NULL) >= 0);
}

ucc_sbgp_t *ucc_topo_get_sbgp(ucc_topo_t *topo, ucc_sbgp_type_t type)
{
    ucc_sbgp_t *s;
    if ((unsigned)type >= UCC_SBGP_LAST) return NULL;
    s = &topo->sbgps[type];
    if (s->status != UCC_SBGP_NOT_INIT) return s;
    s->type = type;
    switch (type) {
    case UCC_SBGP_NUMA: if (!topo->topo->numa_bound) { s->status = UCC_SBGP_NOT_EXISTS; break; } build_local_group(topo, s, same_numa); break;
    case UCC_SBGP_SOCKET: if (!topo->topo->sock_bound) { s->status = UCC_SBGP_NOT_EXISTS; break; } build_local_group(topo, s, same_socket); break;
    case UCC_SBGP_NODE: build_local_group(topo, s, same_node); break;
    case UCC_SBGP_NODE_NVLINK: build_node_nvlink(topo, s); break;
    case UCC_SBGP_NODE_LEADERS: build_leaders(topo, s, same_node, 0, topo->node_leader_rank_id); break;
    case UCC_SBGP_NET: { /* ranks with my local index on every node (rails); needs equal ppn */
        ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, me = topo->set.myrank, my_local = 0, n = 0;
        ucc_rank_t *ranks;
        if (topo->min_ppn != topo->max_ppn) { s->status = UCC_SBGP_NOT_EXISTS; break; }
        ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
        for (ucc_rank_t q = 0; q < me; q++) if (same_node(topo, q, me)) my_local++;
        for (ucc_rank_t r = 0; r < size; r++) { ucc_rank_t l = 0; for (ucc_rank_t q = 0; q < r; q++) if (same_node(topo, q, r)) l++; if (l == my_local) ranks[n++] = r; }
        sbgp_finish(s, topo, ranks, n, 1);
        break; }
    case UCC_SBGP_SOCKET_LEADERS: if (!topo->topo->sock_bound) { s->status = UCC_SBGP_NOT_EXISTS; break; } build_leaders(topo, s, same_socket, 1, 0); break;
    case UCC_SBGP_NUMA_LEADERS: if (!topo->topo->numa_bound) { s->status = UCC_SBGP_NOT_EXISTS; break; } build_leaders(topo, s, same_numa, 1, 0); break;
    case UCC_SBGP_FULL: {
        ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num; ucc_rank_t *ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
        for (ucc_rank_t r = 0; r < size; r++) ranks[r] = r;
        sbgp_finish(s, topo, ranks, size, 1); break; }
    case UCC_SBGP_FULL_HOST_ORDERED: build_full_host_ordered(topo, s); break;
    default: s->status = UCC_SBGP_NOT_EXISTS; break;
    }
    return s;
}

static ucc_status_t all_groups(ucc_topo_t *topo, same_fn_t same, int only_my_node, ucc_sbgp_type_t type, ucc_sbgp_t **out, int *n_out)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, me = topo->set.myrank;
    ucc_sbgp_t *arr = (ucc_sbgp_t *)calloc(size, sizeof(ucc_sbgp_t));
    int n = 0;
    if (!arr) return UCC_ERR_NO_MEMORY;
    for (ucc_rank_t r = 0; r < size; r++) {
        ucc_rank_t first = 1, cnt = 0, *ranks;
        if (only_my_node && !same_node(topo, r, me)) continue;
        for (ucc_rank_t q = 0; q < r; q++) if (same(topo, q, r)) first = 0;
        if (!first) continue;
        ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
        for (ucc_rank_t q = 0; q < size; q++) if (q == r || same(topo, q, r)) ranks[cnt++] = q;
        arr[n].type = type; arr[n].rank_map = ranks; arr[n].group_size = cnt; arr[n].group_rank = UCC_RANK_INVALID;
        for (ucc_rank_t i = 0; i < cnt; i++) if (ranks[i] == me) arr[n].group_rank = i;
        arr[n].map = ucc_ep_map_from_array(&arr[n].rank_map, cnt, size, 0);
        arr[n].status = UCC_SBGP_ENABLED;
        n++;
    }
    *out = arr; *n_out = n;
    return UCC_OK;
}
ucc_status_t ucc_topo_get_all_sockets(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n)
{
    if (!topo->topo->sock_bound) return UCC_ERR_NOT_FOUND;
    if (!topo->all_sockets) UCC_CHECK_RET(all_groups(topo, same_socket, 1, UCC_SBGP_SOCKET, &topo->all_sockets, &topo->n_sockets));
    *sbgps = topo->all_sockets; *n = topo->n_sockets; return UCC_OK;
}
ucc_status_t ucc_topo_get_all_numas(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n)
{
    if (!topo->topo->numa_bound) return UCC_ERR_NOT_FOUND;
    if (!topo->all_numas) UCC_CHECK_RET(all_groups(topo, same_numa, 1, UCC_SBGP_NUMA, &topo->all_numas, &topo->n_numas));
    *sbgps = topo->all_numas; *n = topo->n_numas; return UCC_OK;
}
ucc_status_t ucc_topo_get_all_nodes(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n)
{
    if (!topo->all_nodes) UCC_CHECK_RET(all_groups(topo, same_node, 0, UCC_SBGP_NODE, &topo->all_nodes, &topo->n_nodes_all));
    *sbgps = topo->all_nodes; *n = topo->n_nodes_all; return UCC_OK;
}
int ucc_topo_n_numas(ucc_topo_t *topo)
{
    ucc_sbgp_t *g; int n;
    return ucc_topo_get_all_numas(topo, &g, &n) == UCC_OK ? n : 0;
}
ucc_rank_t ucc_topo_get_node_host_id(ucc_topo_t *topo, ucc_rank_t team_rank)
{
    for (int k = 0; k < topo->n_nodes_all; k++) if (same_node(topo, topo->all_nodes[k].rank_map[0], team_rank)) return (ucc_rank_t)k;
    return UCC_RANK_INVALID;
}
ucc_status_t ucc_topo_get_all_node_nvlinks(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n_out)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num, me = topo->set.myrank, nm = 0, *mem, *comp, *stack;
    ucc_sbgp_t *arr;
    int n = 0;
    if (topo->all_node_nvlinks) { *sbgps = topo->all_node_nvlinks; *n_out = topo->n_node_nvlinks; return UCC_OK; }
    if (ucc_topo_rank_gpu(topo, me, NULL) < 0) return UCC_ERR_NOT_FOUND;
    mem = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size * 3);
    if (!mem) return UCC_ERR_NO_MEMORY;
    comp = mem + size; stack = comp + size;
    for (ucc_rank_t r = 0; r < size; r++) if (same_node(topo, r, me) && ucc_topo_rank_gpu(topo, r, NULL) >= 0) { comp[nm] = UCC_RANK_INVALID; mem[nm++] = r; }
    arr = (ucc_sbgp_t *)calloc(nm, sizeof(ucc_sbgp_t));
    if (!arr) { free(mem); return UCC_ERR_NO_MEMORY; }
    for (ucc_rank_t i = 0; i < nm; i++) { /* flood fill over the NVLink adjacency of the node's GPU members */
        ucc_rank_t top = 0, cnt = 0, *ranks;
        if (comp[i] != UCC_RANK_INVALID) continue;
        comp[i] = (ucc_rank_t)n; stack[top++] = i;
        while (top) {
            ucc_rank_t a = stack[--top];
            for (ucc_rank_t b = 0; b < nm; b++) if (comp[b] == UCC_RANK_INVALID && ucc_topo_nvlink_connected(topo, mem[a], mem[b])) { comp[b] = (ucc_rank_t)n; stack[top++] = b; }
        }
        ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * nm);
        for (ucc_rank_t b = 0; b < nm; b++) if (comp[b] == (ucc_rank_t)n) ranks[cnt++] = mem[b];
        arr[n].type = UCC_SBGP_NODE_NVLINK; arr[n].rank_map = ranks; arr[n].group_size = cnt; arr[n].group_rank = UCC_RANK_INVALID;
        for (ucc_rank_t k = 0; k < cnt; k++) if (ranks[k] == me) arr[n].group_rank = k;
        arr[n].map = ucc_ep_map_from_array(&arr[n].rank_map, cnt, size, 0);
        arr[n].status = UCC_SBGP_ENABLED;
        n++;
    }
    free(mem);
    topo->all_node_nvlinks = arr; topo->n_node_nvlinks = n;
    *sbgps = arr; *n_out = n;
    return UCC_OK;
}
ucc_status_t ucc_topo_get_node_leaders(ucc_topo_t *topo, ucc_rank_t **node_leaders)
{
    ucc_rank_t size = (ucc_rank_t)topo->set.map.ep_num;
    if (!topo->node_leaders) {
        ucc_sbgp_t *nodes; int nn;
        UCC_CHECK_RET(ucc_topo_get_all_nodes(topo, &nodes, &nn));
        topo->node_leaders = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * size);
        if (!topo->node_leaders) return UCC_ERR_NO_MEMORY;
        for (int k = 0; k < nn; k++) {
            ucc_rank_t pick = topo->node_leader_rank_id < nodes[k].group_size ? topo->node_leader_rank_id : nodes[k].group_size - 1;
            for (ucc_rank_t i = 0; i < nodes[k].group_size; i++) topo->node_leaders[nodes[k].rank_map[i]] = nodes[k].rank_map[pick];
        }
        topo->node_leader_rank = topo->node_leaders[topo->set.myrank];
    }
    *node_leaders = topo->node_leaders;
    return UCC_OK;
}

ucc_status_t ucc_topo_init(ucc_subset_t set, ucc_context_topo_t *ctx_topo, ucc_topo_t **topo_p)
{
    ucc_topo_t *t;
    ucc_rank_t size = (ucc_rank_t)set.map.ep_num;
    ucc_sbgp_t *nodes; int nn;
    if (!ctx_topo) return UCC_ERR_INVALID_PARAM;
    t = (ucc_topo_t *)calloc(1, sizeof(*t));
    if (!t) return UCC_ERR_NO_MEMORY;
    t->topo = ctx_topo; t->set = set; t->node_leader_rank_id = 0;
    if (all_groups(t, same_node, 0, UCC_SBGP_NODE, &nodes, &nn) != UCC_OK) { free(t); return UCC_ERR_NO_MEMORY; }
    t->all_nodes = nodes; t->n_nodes_all = nn; t->nnodes = (ucc_rank_t)nn;
    t->min_ppn = UCC_RANK_MAX; t->max_ppn = 0;
    for (int k = 0; k < nn; k++) { if (nodes[k].group_size < t->min_ppn) t->min_ppn = nodes[k].group_size; if (nodes[k].group_size > t->max_ppn) t->max_ppn = nodes[k].group_size; }
    /* socket / numa statistics over the whole set */
    t->min_socket_size = t->min_numa_size = UCC_RANK_MAX;
    for (ucc_rank_t r = 0; r < size; r++) {
        ucc_rank_t ss = 0, ns = 0, first_s = 1, first_n = 1;
        for (ucc_rank_t q = 0; q < size; q++) { if (q == r || same_socket(t, q, r)) { ss++; if (q < r) first_s = 0; } if (q == r || same_numa(t, q, r)) { ns++; if (q < r) first_n = 0; } }
        if (first_s) { if (ss < t->min_socket_size) t->min_socket_size = ss; if (ss > t->max_socket_size) t->max_socket_size = ss; }
        if (first_n) { if (ns < t->min_numa_size) t->min_numa_size = ns; if (ns > t->max_numa_size) t->max_numa_size = ns; }
    }
    for (int k = 0; k < nn; k++) { /* sockets per node */
        ucc_rank_t cnt = 0;
        for (ucc_rank_t i = 0; i < nodes[k].group_size; i++) { int first = 1; for (ucc_rank_t j = 0; j < i; j++) if (same_socket(t, nodes[k].rank_map[i], nodes[k].rank_map[j])) first = 0; cnt += (ucc_rank_t)first; }
        if (cnt > t->max_n_sockets) t->max_n_sockets = cnt;
    }
    *topo_p = t;
    return UCC_OK;
}

static void free_groups(ucc_sbgp_t *arr, int n) { if (!arr) return; for (int i = 0; i < n; i++) free(arr[i].rank_map); free(arr); }
void ucc_topo_cleanup(ucc_topo_t *t)
{
    if (!t) return;
    for (int i = 0; i < UCC_SBGP_LAST; i++) free(t->sbgps[i].rank_map);
    free_groups(t->all_sockets, t->n_sockets); free_groups(t->all_numas, t->n_numas); free_groups(t->all_nodes, t->n_nodes_all); free_groups(t->all_node_nvlinks, t->n_node_nvlinks);
    free(t->node_leaders); free(t);
}
void ucc_sbgp_print(const ucc_sbgp_t *s)
{
    char buf[1024]; size_t o = 0;
    if (s->status != UCC_SBGP_ENABLED) return;
    for (ucc_rank_t i = 0; i < s->group_size && o < sizeof(buf); i++) o += snprintf(buf + o, sizeof(buf) - o, "%u ", s->rank_map[i]);
    ucc_print("sbgp %s: group_size %u group_rank %u ranks: %s", ucc_sbgp_str(s->type), s->group_size, s->group_rank, buf);
}
