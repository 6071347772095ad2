/* shim: see libacars.h */
#ifndef ORACLE_SHIM_LA_HASH_H
#define ORACLE_SHIM_LA_HASH_H
#include <stdint.h>
#include <stdbool.h>
typedef uint32_t (la_hash_func)(void const *key);
typedef bool (la_hash_compare_func)(void const *key1, void const *key2);
typedef void (la_hash_key_destroy_func)(void *key);
#endif
