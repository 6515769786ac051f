/* mplb_internal.h — the two host helpers the translation units of libmplb.so share (not part of the ABI). */
#ifndef MPLB_INTERNAL_H
#define MPLB_INTERNAL_H
#if defined(__GNUC__)
#define MPLB_HIDDEN __attribute__((visibility("hidden")))
#else
#define MPLB_HIDDEN
#endif
/* records `msg` for mplb_last_error() and returns `code` */
MPLB_HIDDEN int mplb_internal_fail(int code, const char *msg);
/* adds to the counter behind mplb_launch_count() */
MPLB_HIDDEN void mplb_internal_count_launches(int n);
#endif
