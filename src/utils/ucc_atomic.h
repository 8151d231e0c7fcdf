/* Atomics, spinlocks (plain + recursive) and a thread-mode aware lock. */
#ifndef UCC_ATOMIC_H_
#define UCC_ATOMIC_H_
#include "ucc_compiler_def.h"
#include <pthread.h>
#include <sched.h>

#define ucc_atomic_add32(_p, _v)  __atomic_add_fetch((_p), (_v), __ATOMIC_ACQ_REL)
#define ucc_atomic_fadd32(_p, _v) __atomic_fetch_add((_p), (_v), __ATOMIC_ACQ_REL)
#define ucc_atomic_add64(_p, _v)  __atomic_add_fetch((_p), (_v), __ATOMIC_ACQ_REL)
#define ucc_atomic_fadd64(_p, _v) __atomic_fetch_add((_p), (_v), __ATOMIC_ACQ_REL)
#define ucc_atomic_cswap64(_p, _c, _s) ({ uint64_t _e = (_c); __atomic_compare_exchange_n((_p), &_e, (_s), 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); _e; })
#define ucc_atomic_cswap32(_p, _c, _s) ({ uint32_t _e = (_c); __atomic_compare_exchange_n((_p), &_e, (_s), 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); _e; })
#define ucc_atomic_bool_cswap64(_p, _c, _s) ({ uint64_t _e = (_c); __atomic_compare_exchange_n((_p), &_e, (_s), 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); })
#define ucc_load_acquire(_p)      __atomic_load_n((_p), __ATOMIC_ACQUIRE)
#define ucc_store_release(_p, _v) __atomic_store_n((_p), (_v), __ATOMIC_RELEASE)
#define ucc_memory_cpu_fence()       __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define ucc_memory_cpu_store_fence() __atomic_thread_fence(__ATOMIC_RELEASE)
#define ucc_memory_cpu_load_fence()  __atomic_thread_fence(__ATOMIC_ACQUIRE)
#if defined(__x86_64__)
#define ucc_cpu_relax() __builtin_ia32_pause()
#else
#define ucc_cpu_relax() sched_yield()
#endif

typedef struct ucc_spinlock { volatile int v; } ucc_spinlock_t;
static inline void ucc_spinlock_init(ucc_spinlock_t *l) { l->v = 0; }
static inline void ucc_spinlock_destroy(ucc_spinlock_t *l) { (void)l; }
static inline int  ucc_spin_trylock(ucc_spinlock_t *l) { return !__atomic_exchange_n(&l->v, 1, __ATOMIC_ACQUIRE); }
static inline void ucc_spin_lock(ucc_spinlock_t *l)
{ while (!ucc_spin_trylock(l)) { while (__atomic_load_n(&l->v, __ATOMIC_RELAXED)) ucc_cpu_relax(); } }
static inline void ucc_spin_unlock(ucc_spinlock_t *l) { __atomic_store_n(&l->v, 0, __ATOMIC_RELEASE); }

typedef struct ucc_recursive_spinlock { ucc_spinlock_t l; pthread_t owner; int count; } ucc_recursive_spinlock_t;
static inline void ucc_recursive_spinlock_init(ucc_recursive_spinlock_t *l) { ucc_spinlock_init(&l->l); l->owner = 0; l->count = 0; }
static inline void ucc_recursive_spin_lock(ucc_recursive_spinlock_t *l)
{
    pthread_t self = pthread_self();
    /* owner is only ever equal to `self` if this thread stored it, so a relaxed read is enough */
    if (pthread_equal(__atomic_load_n(&l->owner, __ATOMIC_RELAXED), self)) { l->count++; return; }
    ucc_spin_lock(&l->l); __atomic_store_n(&l->owner, self, __ATOMIC_RELAXED); l->count = 1;
}
static inline void ucc_recursive_spin_unlock(ucc_recursive_spinlock_t *l)
{ if (--l->count == 0) { __atomic_store_n(&l->owner, (pthread_t)0, __ATOMIC_RELAXED); ucc_spin_unlock(&l->l); } }

/* lock that degenerates to a no-op for UCC_THREAD_SINGLE */
typedef struct ucc_tm_lock { int enabled; ucc_spinlock_t l; } ucc_tm_lock_t;
static inline void ucc_tm_lock_init(ucc_tm_lock_t *l, int enabled) { l->enabled = enabled; ucc_spinlock_init(&l->l); }
static inline void ucc_tm_lock(ucc_tm_lock_t *l) { if (l->enabled) ucc_spin_lock(&l->l); }
static inline void ucc_tm_unlock(ucc_tm_lock_t *l) { if (l->enabled) ucc_spin_unlock(&l->l); }
#endif
