/* shim: see libacars.h. la_list must be complete: reference src/decode.c:411 reads p->data. */
#ifndef ORACLE_SHIM_LA_LIST_H
#define ORACLE_SHIM_LA_LIST_H
typedef struct la_list { void *data; struct la_list *next; } la_list;
la_list *la_list_next(la_list const *l);
void la_list_foreach(la_list *l, void (*cb)(void *data, void *ctx), void *ctx);
#endif
