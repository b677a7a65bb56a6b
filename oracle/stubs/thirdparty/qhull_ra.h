/* Stand-in for qhull's reentrant API (not in this image): convex hulls of meshes are outside the
 * mj_inverse path we check. Types and iteration macros exist so the reference's mesh compiler
 * builds unchanged; calling qh_qhull aborts. */
#ifndef ORACLE_STUB_QHULL_RA_H_
#define ORACLE_STUB_QHULL_RA_H_
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
typedef double coordT;
typedef coordT pointT;
typedef unsigned int boolT;
#define False 0
#define True 1
#define qh_ALL True
typedef struct setT { int maxsize; void* e[1]; } setT;
typedef struct vertexT vertexT;
typedef struct facetT facetT;
struct vertexT { vertexT* next; vertexT* previous; pointT* point; setT* neighbors; };
struct facetT { facetT* next; facetT* previous; setT* vertices; unsigned toporient : 1; };
typedef struct qhT {
  jmp_buf errexit;
  boolT NOerrexit;
  int num_vertices, num_facets;
  vertexT* vertex_list;
  facetT* facet_list;
} qhT;
#define FORALLvertices for (vertex = qh->vertex_list; vertex && vertex->next; vertex = vertex->next)
#define FORALLfacets for (facet = qh->facet_list; facet && facet->next; facet = facet->next)
#define FOREACHsetelement_(type, set, variable) \
  if (((variable = NULL), set)) \
    for (variable##p = (type**)&((set)->e[0]); (variable = *variable##p++);)
static inline void qh_unavailable_(void) {
  fprintf(stderr, "oracle/_ref: qhull is not available in this build\n");
  abort();
}
static inline void qh_zero(qhT* qh, FILE* f) { (void)qh; (void)f; }
static inline void qh_init_A(qhT* qh, FILE* a, FILE* b, FILE* c, int argc, char** argv) {
  (void)qh; (void)a; (void)b; (void)c; (void)argc; (void)argv;
}
static inline void qh_initflags(qhT* qh, char* s) { (void)qh; (void)s; }
static inline void qh_init_B(qhT* qh, coordT* p, int n, int d, boolT m) { (void)qh; (void)p; (void)n; (void)d; (void)m; }
static inline void qh_qhull(qhT* qh) { (void)qh; qh_unavailable_(); }
static inline void qh_triangulate(qhT* qh) { (void)qh; }
static inline void qh_vertexneighbors(qhT* qh) { (void)qh; }
static inline int qh_pointid(qhT* qh, pointT* p) { (void)qh; (void)p; return -1; }
static inline void qh_freeqhull(qhT* qh, boolT all) { (void)qh; (void)all; }
static inline void qh_memfreeshort(qhT* qh, int* a, int* b) { (void)qh; *a = 0; *b = 0; }
#endif
