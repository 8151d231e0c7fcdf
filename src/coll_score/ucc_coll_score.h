/* Algorithm selection by score.
 *
 * Every CL/TL team publishes, per (collective, memory type), a sorted list
 * of message ranges [start,end) with a score and an init function.  Lists
 * are merged (higher score wins, the loser is kept as a fallback), can be
 * overridden with a TUNE string, and are finally frozen into a map used
 * by ucc_collective_init.
 *
 * Behavioural contract: reference coll_score/ucc_coll_score.h:46-173 and
 * the TUNE grammar of components/base/ucc_base_iface.c:32-52.  The
 * implementation uses a boundary sweep over elementary intervals instead
 * of the reference's pairwise list surgery.
 */
#ifndef UCC_COLL_SCORE_H_
#define UCC_COLL_SCORE_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_compiler_def.h"
#include "utils/ucc_list.h"
#include "utils/ucc_coll_utils.h"
#include "schedule/ucc_schedule.h"

#define UCC_SCORE_MAX_STR "inf"

typedef struct ucc_coll_entry {
    ucc_list_link_t         list_elem;
    ucc_score_t             score;
    ucc_base_coll_init_fn_t init;
    ucc_base_team_t        *team;
} ucc_coll_entry_t;

typedef struct ucc_msg_range {
    ucc_coll_entry_t super;
    ucc_list_link_t  fallback; /* list of ucc_coll_entry_t, best first */
    size_t           start, end; /* [start, end); end == UCC_MSG_MAX is open */
} ucc_msg_range_t;

typedef struct ucc_coll_score {
    ucc_list_link_t scores[UCC_COLL_TYPE_NUM][UCC_MEMORY_TYPE_LAST];
} ucc_coll_score_t;

typedef struct ucc_score_map ucc_score_map_t;

/* alg id/name -> init fn, provided by each CL/TL for "@alg" tokens */
typedef ucc_status_t (*ucc_alg_id_to_init_fn_t)(int alg_id, const char *alg_id_str, ucc_coll_type_t coll_type,
                                                ucc_memory_type_t mem_type, ucc_base_coll_init_fn_t *init);

typedef struct ucc_coll_score_team_info {
    ucc_score_t             default_score;
    ucc_rank_t              size;
    uint64_t                supported_colls;
    ucc_memory_type_t      *supported_mem_types;
    int                     num_mem_types;
    ucc_base_coll_init_fn_t init;
    ucc_alg_id_to_init_fn_t alg_fn;
} ucc_coll_score_team_info_t;

ucc_status_t ucc_coll_score_alloc(ucc_coll_score_t **score);
void         ucc_coll_score_free(ucc_coll_score_t *score);
ucc_status_t ucc_coll_score_add_range(ucc_coll_score_t *score, ucc_coll_type_t coll_type, ucc_memory_type_t mem_type,
                                      size_t start, size_t end, ucc_score_t msg_score, ucc_base_coll_init_fn_t init,
                                      ucc_base_team_t *team);
/* out = max(score1, score2) per interval; loser appended to the winner's fallbacks */
ucc_status_t ucc_coll_score_merge(ucc_coll_score_t *score1, ucc_coll_score_t *score2, ucc_coll_score_t **rst, int free_inputs);
ucc_status_t ucc_coll_score_merge_in(ucc_coll_score_t **dst, ucc_coll_score_t *src); /* dst = merge(dst, src), frees both inputs */
ucc_status_t ucc_coll_score_dup(const ucc_coll_score_t *in, ucc_coll_score_t **out);
void         ucc_coll_score_set(ucc_coll_score_t *score, ucc_score_t value);
/* build [0,inf) ranges with default score for the listed colls / mem types */
ucc_status_t ucc_coll_score_build_default(ucc_base_team_t *team, ucc_score_t default_score, ucc_base_coll_init_fn_t default_init,
                                          uint64_t coll_types, ucc_memory_type_t *mem_types, int mt_n, ucc_coll_score_t **score);
/* parse a TUNE string into a (partial) score: entries may carry UCC_SCORE_INVALID / NULL init = "keep" */
ucc_status_t ucc_coll_score_alloc_from_str(const char *str, ucc_coll_score_t **score, ucc_rank_t team_size,
                                           ucc_base_coll_init_fn_t init, ucc_base_team_t *team, ucc_alg_id_to_init_fn_t alg_fn);
/* apply a partial score on top of `score` */
ucc_status_t ucc_coll_score_update(ucc_coll_score_t *score, ucc_coll_score_t *update, ucc_score_t default_score,
                                   ucc_memory_type_t *mtypes, int mt_n, uint64_t colls);
ucc_status_t ucc_coll_score_update_from_str(const char *str, const ucc_coll_score_team_info_t *info, ucc_base_team_t *team,
                                            ucc_coll_score_t *score);

ucc_status_t ucc_coll_score_build_map(ucc_coll_score_t *score, ucc_score_map_t **map); /* takes ownership of score */
void         ucc_coll_score_free_map(ucc_score_map_t *map);
/* select by (coll, mem type, msg size) and run init; walks fallbacks on NOT_SUPPORTED / NOT_IMPLEMENTED */
ucc_status_t ucc_coll_init(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_coll_task_t **task);
ucc_status_t ucc_coll_init_as(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_rank_t rank, ucc_rank_t size, ucc_coll_task_t **task);
ucc_status_t ucc_coll_score_map_lookup(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_msg_range_t **range);
void         ucc_coll_score_map_print_info(const ucc_score_map_t *map, int verbosity);
void         ucc_coll_score_map_str(const ucc_score_map_t *map, char *buf, size_t len);
/* lets components give human names to init functions for the map dump */
typedef const char *(*ucc_coll_score_name_fn_t)(ucc_base_coll_init_fn_t init, ucc_base_team_t *team);
void         ucc_coll_score_set_name_resolver(ucc_coll_score_name_fn_t fn);
#endif
