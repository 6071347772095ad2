/* shim: see libacars.h */
#ifndef ORACLE_SHIM_LA_REASM_H
#define ORACLE_SHIM_LA_REASM_H
typedef struct la_reasm_ctx la_reasm_ctx;
la_reasm_ctx *la_reasm_ctx_new(void);
#endif
