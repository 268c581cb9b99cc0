/* Stand-in for <omp.h> (the toolchain of this image ships no libgomp): the reference's detect path is serial; the
 * OpenMP pragmas of its trainer are ignored without -fopenmp and these calls report one thread.  Test infrastructure. */
#ifndef JDA_CVSHIM_OMP_H_
#define JDA_CVSHIM_OMP_H_
typedef int omp_lock_t;
static inline int omp_get_thread_num(void) { return 0; }
static inline int omp_get_max_threads(void) { return 1; }
static inline int omp_get_num_threads(void) { return 1; }
static inline void omp_set_num_threads(int n) { (void)n; }
static inline void omp_init_lock(omp_lock_t *l) { *l = 0; }
static inline void omp_destroy_lock(omp_lock_t *l) { (void)l; }
static inline void omp_set_lock(omp_lock_t *l) { *l = 1; }
static inline void omp_unset_lock(omp_lock_t *l) { *l = 0; }
#endif
