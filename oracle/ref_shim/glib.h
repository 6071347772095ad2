/* oracle/ref_shim/glib.h — minimal stand-in for <glib.h> (glib is not installed here).
 * Only the GAsyncQueue calls used by reference src/decode.c:53,170,393,456,530 are declared;
 * oracle/ref_harness.c implements them (push == capture the frame). TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GLIB_H
#define ORACLE_SHIM_GLIB_H
typedef struct shim_async_queue GAsyncQueue;
GAsyncQueue *g_async_queue_new(void);
void g_async_queue_push(GAsyncQueue *q, void *item);
void *g_async_queue_pop(GAsyncQueue *q);
int g_async_queue_length(GAsyncQueue *q);
#endif
