/* Team object and creation state machine (contract: reference core/ucc_team.{h,c}). */
#ifndef UCC_TEAM_H_
#define UCC_TEAM_H_
#include <ucc/api/ucc.h>
#include "ucc_context.h"
#include "utils/ucc_math.h"

typedef struct ucc_service_coll_req ucc_service_coll_req_t;
typedef struct ucc_topo ucc_topo_t;
typedef enum {
    UCC_TEAM_ADDR_EXCHANGE, UCC_TEAM_SERVICE_TEAM, UCC_TEAM_ALLOC_ID, UCC_TEAM_CL_CREATE, UCC_TEAM_ACTIVE
} ucc_team_state_t;

typedef struct ucc_team {
    ucc_status_t            status;
    ucc_team_state_t        state;
    ucc_context_t         **contexts;
    uint32_t                num_contexts;
    ucc_base_team_params_t  bp;
    ucc_team_oob_coll_t     oob;
    ucc_cl_team_t         **cl_teams;
    int                     n_cl_teams;
    int                     last_team_create_posted;
    uint16_t                id;        /* bit 15 set: user supplied id */
    ucc_rank_t              rank, size;
    ucc_tl_team_t          *service_team;
    int                     service_team_owned;
    ucc_service_coll_req_t *sreq;
    ucc_addr_storage_t      addr_storage; /* only when the context has none */
    ucc_rank_t             *ctx_ranks;
    void                   *oob_req;
    ucc_ep_map_t            ctx_map;   /* team rank -> context rank */
    ucc_topo_t             *topo;
    struct ucc_context_topo *own_ctx_topo; /* only when the context has no address storage */
    ucc_score_map_t        *score_map;
    uint32_t                seq_num;   /* collectives initialised on this team */
    uint64_t               *id_bitmap_work; /* scratch of the id allocation */
    ucc_rank_t              ctx_rank_send;
} ucc_team_t;

#define UCC_TEAM_ID_IS_EXTERNAL(_t) ((_t)->id & 0x8000)
#define UCC_TEAM_ID_EXTERNAL_BIT 0x8000
#define UCC_TEAM_ID_MAX 0x7fff

static inline ucc_rank_t ucc_team_rank_(const ucc_team_t *t) { return t->rank; }
static inline ucc_rank_t ucc_team_size_(const ucc_team_t *t) { return t->size; }
/* team rank -> context rank */
static inline ucc_rank_t ucc_get_ctx_rank(const ucc_team_t *team, ucc_rank_t team_rank) { return ucc_ep_map_eval(team->ctx_map, team_rank); }
/* address storage holding the endpoints of this team (context-wide or team-private) */
static inline ucc_addr_storage_t *ucc_team_addr_storage(ucc_team_t *team, int *indexed_by_ctx_rank)
{
    if (team->addr_storage.storage) { *indexed_by_ctx_rank = 0; return &team->addr_storage; }
    *indexed_by_ctx_rank = 1;
    return &team->contexts[0]->addr_storage;
}
/* component address blob of a team member; NULL if the peer does not run that component */
static inline void *ucc_get_team_ep_addr(ucc_context_t *ctx, ucc_team_t *team, ucc_rank_t rank, unsigned long component_id, uint32_t *len)
{
    int by_ctx; ucc_addr_storage_t *s = ucc_team_addr_storage(team, &by_ctx);
    (void)ctx;
    return ucc_addr_storage_component_addr(s, by_ctx ? ucc_get_ctx_rank(team, rank) : rank, component_id, len);
}
static inline ucc_context_addr_header_t *ucc_get_team_ep_header(ucc_context_t *ctx, ucc_team_t *team, ucc_rank_t rank)
{
    int by_ctx; ucc_addr_storage_t *s = ucc_team_addr_storage(team, &by_ctx);
    (void)ctx;
    return UCC_ADDR_STORAGE_RANK_HEADER(s, by_ctx ? ucc_get_ctx_rank(team, rank) : rank);
}
void ucc_team_release_id(ucc_team_t *team);
#endif
