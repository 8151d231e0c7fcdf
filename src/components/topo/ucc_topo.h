/* Topology: context-wide process table, per-team view and on-demand
 * sub-groups (sbgp).  Capabilities follow reference topo/ucc_topo.h:93-345
 * and topo/ucc_sbgp.h:10-41; data layout and algorithms are new (single
 * pass bucketing over (host, socket, numa) keys). */
#ifndef UCC_TOPO_H_
#define UCC_TOPO_H_
#include "utils/ucc_proc_info.h"
#include "utils/ucc_coll_utils.h"

typedef struct ucc_addr_storage ucc_addr_storage_t;
typedef struct ucc_team ucc_team_t;

typedef struct ucc_context_topo {
    ucc_proc_info_t *procs;      /* indexed by context rank */
    ucc_host_info_t *hosts;      /* host_info of each context rank */
    ucc_rank_t       n_procs;
    ucc_rank_t       nnodes;
    ucc_rank_t       min_ppn, max_ppn;
    ucc_rank_t       max_n_sockets; /* max sockets used on a node */
    ucc_rank_t       min_socket_size, max_socket_size;
    ucc_rank_t       min_numa_size, max_numa_size;
    int              sock_bound, numa_bound; /* every proc reports a socket / numa */
} ucc_context_topo_t;

typedef enum {
    UCC_SBGP_NUMA, UCC_SBGP_SOCKET, UCC_SBGP_NODE, UCC_SBGP_NODE_NVLINK, UCC_SBGP_NODE_LEADERS, UCC_SBGP_NET,
    UCC_SBGP_SOCKET_LEADERS, UCC_SBGP_NUMA_LEADERS, UCC_SBGP_FULL, UCC_SBGP_FULL_HOST_ORDERED, UCC_SBGP_LAST
} ucc_sbgp_type_t;
typedef enum { UCC_SBGP_NOT_INIT, UCC_SBGP_DISABLED, UCC_SBGP_ENABLED, UCC_SBGP_NOT_EXISTS } ucc_sbgp_status_t;

typedef struct ucc_sbgp {
    ucc_sbgp_type_t   type;
    ucc_sbgp_status_t status;
    ucc_rank_t        group_size;
    ucc_rank_t        group_rank;
    ucc_rank_t       *rank_map;   /* sbgp rank -> team(set) rank */
    ucc_ep_map_t      map;
    ucc_team_t       *team;
} ucc_sbgp_t;

typedef struct ucc_topo {
    ucc_context_topo_t *topo;
    ucc_subset_t        set;          /* team rank -> ctx rank, + my rank */
    ucc_sbgp_t          sbgps[UCC_SBGP_LAST];
    ucc_sbgp_t         *all_sockets;  /* every socket group of my node */
    int                 n_sockets;
    ucc_sbgp_t         *all_numas;
    int                 n_numas;
    ucc_sbgp_t         *all_nodes;    /* every node group of the team */
    int                 n_nodes_all;
    ucc_sbgp_t         *all_node_nvlinks; /* NVLink islands (connected components) among the GPU members of my node */
    int                 n_node_nvlinks;
    ucc_rank_t          node_leader_rank_id; /* which local rank acts as leader (default 0) */
    ucc_rank_t          node_leader_rank;
    ucc_rank_t         *node_leaders; /* team rank -> team rank of its node leader */
    ucc_rank_t          min_ppn, max_ppn, nnodes;
    ucc_rank_t          min_socket_size, max_socket_size, max_n_sockets;
    ucc_rank_t          min_numa_size, max_numa_size;
} ucc_topo_t;

ucc_status_t ucc_context_topo_init(ucc_addr_storage_t *storage, ucc_context_topo_t **topo);
void         ucc_context_topo_cleanup(ucc_context_topo_t *topo);
ucc_status_t ucc_topo_init(ucc_subset_t set, ucc_context_topo_t *ctx_topo, ucc_topo_t **topo);
void         ucc_topo_cleanup(ucc_topo_t *topo);
ucc_sbgp_t  *ucc_topo_get_sbgp(ucc_topo_t *topo, ucc_sbgp_type_t type);
ucc_status_t ucc_topo_get_all_sockets(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n_sbgps);
ucc_status_t ucc_topo_get_all_numas(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n_sbgps);
ucc_status_t ucc_topo_get_all_nodes(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n_sbgps);
ucc_status_t ucc_topo_get_node_leaders(ucc_topo_t *topo, ucc_rank_t **node_leaders);
const char  *ucc_sbgp_str(ucc_sbgp_type_t type);
void         ucc_sbgp_print(const ucc_sbgp_t *sbgp);

static inline ucc_rank_t ucc_topo_nnodes(const ucc_topo_t *t) { return t->nnodes; }
static inline ucc_rank_t ucc_topo_min_ppn(const ucc_topo_t *t) { return t->min_ppn; }
static inline ucc_rank_t ucc_topo_max_ppn(const ucc_topo_t *t) { return t->max_ppn; }
static inline int ucc_topo_is_single_node(const ucc_topo_t *t) { return t->nnodes == 1; }
static inline int ucc_topo_isoppn(const ucc_topo_t *t) { return t->min_ppn == t->max_ppn; }
static inline ucc_rank_t ucc_topo_n_sockets(const ucc_topo_t *t) { return t->max_n_sockets; }
/* further statistics of the reference's inline family (topo/ucc_topo.h:148-282) */
static inline int ucc_topo_is_single_ppn(const ucc_topo_t *t) { return t->max_ppn == 1; }
static inline ucc_rank_t ucc_topo_min_socket_size(const ucc_topo_t *t) { return t->min_socket_size; }
static inline ucc_rank_t ucc_topo_max_socket_size(const ucc_topo_t *t) { return t->max_socket_size; }
static inline ucc_rank_t ucc_topo_min_numa_size(const ucc_topo_t *t) { return t->min_numa_size; }
static inline ucc_rank_t ucc_topo_max_numa_size(const ucc_topo_t *t) { return t->max_numa_size; }
int          ucc_topo_n_numas(ucc_topo_t *topo);                 /* numa domains used on my node, 0 when processes are not bound */
/* dense id (0 .. nnodes-1, order of first appearance in the team) of the node a team rank lives on */
ucc_rank_t   ucc_topo_get_node_host_id(ucc_topo_t *topo, ucc_rank_t team_rank);
/* NVLink islands of my node: one group per connected component of the NVLink graph over the node's GPU members
 * (reference ucc_sbgp_create_all_node_nvlinks, topo/ucc_sbgp.h:90); UCC_ERR_NOT_FOUND without device information */
ucc_status_t ucc_topo_get_all_node_nvlinks(ucc_topo_t *topo, ucc_sbgp_t **sbgps, int *n_sbgps);
/* a sub-group as a subset of the team it was cut from (for sub-team creation and service collectives) */
static inline ucc_subset_t ucc_sbgp_to_subset(const ucc_sbgp_t *sbgp) { ucc_subset_t s; s.map = sbgp->map; s.myrank = sbgp->group_rank; return s; }
/* device predicates */
int ucc_topo_has_device_info(const ucc_topo_t *topo);           /* every member reports a GPU */
int ucc_topo_is_nvlink_fully_connected(const ucc_topo_t *topo); /* every pair of members is NVLink reachable */
int ucc_topo_is_single_nvlink_domain(const ucc_topo_t *topo);   /* same node or same fabric clique */
int ucc_topo_nvlink_connected(const ucc_topo_t *topo, ucc_rank_t r1, ucc_rank_t r2);
int ucc_topo_rank_gpu(const ucc_topo_t *topo, ucc_rank_t team_rank, ucc_gpu_info_t **gpu); /* gpu index or -1 */
#endif
