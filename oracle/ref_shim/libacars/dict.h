/* shim: see libacars.h */
#ifndef ORACLE_SHIM_LA_DICT_H
#define ORACLE_SHIM_LA_DICT_H
typedef struct la_dict la_dict;
#endif
