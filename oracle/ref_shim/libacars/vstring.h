/* shim: see libacars.h */
#ifndef ORACLE_SHIM_LA_VSTRING_H
#define ORACLE_SHIM_LA_VSTRING_H
typedef struct la_vstring la_vstring;
#endif
