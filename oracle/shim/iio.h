/*
 * TEST INFRASTRUCTURE ONLY. Stand-in for <iio.h> (libiio, an un-vendored dependency of the reference's PlutoSDR plugin):
 * the reference reaches libiio only through function pointers it looks up at run time (src/sdr/iio_lib.c) or that its tests
 * replace with mocks (test/iio_lib_mock.c), so the opaque handle types are all that its sources need at compile time.
 */
#ifndef SDRM_IIO_SHIM_H
#define SDRM_IIO_SHIM_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

struct iio_context;
struct iio_device;
struct iio_channel;
struct iio_buffer;
struct iio_scan_context;
struct iio_context_info;

#endif
