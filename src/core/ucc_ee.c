#include "ucc_ee.h"
#include "ucc_team.h"
#include "utils/ucc_log.h"

UCC_EXPORT ucc_status_t ucc_ee_create(ucc_team_h team, const ucc_ee_params_t *params, ucc_ee_h *ee_p)
{
    ucc_ee_t *ee;
    if (!team || !params || !ee_p) return UCC_ERR_INVALID_PARAM;
    ee = (ucc_ee_t *)calloc(1, sizeof(*ee));
    if (!ee) { ucc_error("failed to allocate %zd bytes for ee", sizeof(*ee)); return UCC_ERR_NO_MEMORY; }
    ee->team = team; ee->ee_type = params->ee_type; ee->ee_context_size = params->ee_context_size; ee->ee_context = params->ee_context;
    ucc_spinlock_init(&ee->lock);
    ucc_queue_head_init(&ee->event_in_queue); ucc_queue_head_init(&ee->event_out_queue);
    *ee_p = ee;
    return UCC_OK;
}
static void drain(ucc_queue_head_t *q)
{ ucc_queue_elem_t *e; while ((e = ucc_queue_pull(q))) free(ucc_container_of(e, ucc_event_desc_t, queue)); }
UCC_EXPORT ucc_status_t ucc_ee_destroy(ucc_ee_h ee)
{ if (!ee) return UCC_ERR_INVALID_PARAM; drain(&ee->event_in_queue); drain(&ee->event_out_queue); free(ee); return UCC_OK; }

ucc_status_t ucc_ee_get_event_internal(ucc_ee_h ee, ucc_ev_t **ev, ucc_queue_head_t *queue)
{
    ucc_queue_elem_t *e;
    ucc_spin_lock(&ee->lock);
    e = ucc_queue_pull(queue);
    ucc_spin_unlock(&ee->lock);
    if (!e) return UCC_ERR_NOT_FOUND;
    *ev = &ucc_container_of(e, ucc_event_desc_t, queue)->ev;
    return UCC_OK;
}
ucc_status_t ucc_ee_set_event_internal(ucc_ee_h ee, ucc_ev_t *ev, ucc_queue_head_t *queue)
{
    ucc_event_desc_t *d = (ucc_event_desc_t *)malloc(sizeof(*d));
    if (!d) return UCC_ERR_NO_MEMORY;
    d->ev = *ev;
    ucc_spin_lock(&ee->lock); ucc_queue_push(queue, &d->queue); ucc_spin_unlock(&ee->lock);
    return UCC_OK;
}
UCC_EXPORT ucc_status_t ucc_ee_get_event(ucc_ee_h ee, ucc_ev_t **ev) { return ucc_ee_get_event_internal(ee, ev, &ee->event_out_queue); }
UCC_EXPORT ucc_status_t ucc_ee_ack_event(ucc_ee_h ee, ucc_ev_t *ev)
{ (void)ee; if (!ev) return UCC_ERR_INVALID_PARAM; free(ucc_container_of(ev, ucc_event_desc_t, ev)); return UCC_OK; }
UCC_EXPORT ucc_status_t ucc_ee_set_event(ucc_ee_h ee, ucc_ev_t *ev) { return ucc_ee_set_event_internal(ee, ev, &ee->event_in_queue); }
UCC_EXPORT ucc_status_t ucc_ee_wait(ucc_ee_h ee, ucc_ev_t *ev)
{
    /* blocks until an event of the requested type shows up in the out queue */
    ucc_ev_t *got = NULL;
    ucc_team_t *team = (ucc_team_t *)ee->team;
    for (;;) {
        if (ucc_ee_get_event_internal(ee, &got, &ee->event_out_queue) == UCC_OK) {
            int match = !ev || got->ev_type == ev->ev_type;
            if (ev) *ev = *got;
            ucc_ee_ack_event(ee, got);
            if (match) return UCC_OK;
        }
        ucc_context_progress(team->contexts[0]);
    }
}
