/* oracle/ref_shim/config.h — stand-in for the CMake-generated config.h of the
 * reference (template: reference src/config.h.in). TEST INFRASTRUCTURE ONLY:
 * lets the unmodified reference hot-path sources compile in place. */
#ifndef _CONFIG_H
#define _CONFIG_H
#define HAVE_PTHREAD_BARRIERS
#define SINCOSF sincosf
#endif
