/* oracle/ref_shim/libacars/libacars.h — opaque stand-ins for the libacars types that the
 * reference headers merely mention (libacars is not installed here). TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_LIBACARS_H
#define ORACLE_SHIM_LIBACARS_H
typedef struct la_proto_node la_proto_node;
typedef struct { int shim_unused; } la_type_descriptor;
typedef enum { LA_MSG_DIR_UNKNOWN, LA_MSG_DIR_GND2AIR, LA_MSG_DIR_AIR2GND } la_msg_dir;
void la_proto_tree_destroy(la_proto_node *root);
#endif
