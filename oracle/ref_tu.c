/*
 * ref_tu.c -- the one translation unit that turns the UNMODIFIED reference
 * header into a shared object (oracle/_ref/libqoi_ref.so).  The header is
 * compiled from where it lies under /root/reference (include path given by
 * oracle/Makefile); no reference source is copied into this repository.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
 */
#define QOI_IMPLEMENTATION
#include "qoi.h"
