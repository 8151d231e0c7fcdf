#ifndef UCC_TIME_H_
#define UCC_TIME_H_
#include <time.h>
#include <sys/time.h>
static inline double ucc_get_time(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
static inline double ucc_get_wall_time(void)
{ struct timeval tv; gettimeofday(&tv, NULL); return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec; }
#endif
