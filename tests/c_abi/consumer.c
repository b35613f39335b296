/* A plain-C (C99) consumer of include/b200ctl.h: what a non-Python host binding sees.  Built and run by
 * tests/test_abi.py::test_plain_c_consumer with gcc -std=c99 -pedantic -Werror; needs no GPU: it checks the version, that
 * argument errors come back as negative codes with a message (never a crash, never a CUDA call), and the struct layouts a
 * foreign binding has to reproduce. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "b200ctl.h"

static int fails = 0;
#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) { printf("FAIL line %d: %s\n", __LINE__, #cond); ++fails; } \
  } while (0)

int main(void) {
  int64_t shape[2] = {4, 3};
  float host_data[12];
  DLTensor t;
  int rc;

  CHECK(b200ctl_version() == B200CTL_VERSION);
  CHECK(sizeof(DLTensor) == 48 && offsetof(DLTensor, shape) == 24 && offsetof(DLTensor, byte_offset) == 40);
  CHECK(sizeof(b200ctl_servo_params) == 96);
  CHECK(B200CTL_STATS_LEN == 8 && B200CTL_STAT_N_ENV == 0 && B200CTL_STAT_N_NONFINITE == 4);

  /* NULL tensors: E_NULL and a message */
  rc = b200ctl_cclvf(NULL, NULL, 1.0, 1.0, NULL, NULL);
  CHECK(rc == B200CTL_E_NULL);
  CHECK(strstr(b200ctl_last_error(), "NULL") != NULL);

  /* a HOST tensor handed to a device entry point is refused by validation (no CPU path behind this ABI) */
  memset(host_data, 0, sizeof host_data);
  t.data = host_data;
  t.device.device_type = kDLCPU;
  t.device.device_id = 0;
  t.ndim = 2;
  t.dtype.code = kDLFloat;
  t.dtype.bits = 32;
  t.dtype.lanes = 1;
  t.shape = shape;
  t.strides = NULL;
  t.byte_offset = 0;
  rc = b200ctl_cclvf(&t, &t, 1.0, 1.0, &t, NULL);
  CHECK(rc < 0);
  CHECK(b200ctl_last_error()[0] != '\0');
  rc = b200ctl_orientation_error(&t, &t, &t, NULL);
  CHECK(rc < 0);

  if (fails == 0) printf("c_abi consumer OK (version %d)\n", b200ctl_version());
  return fails;
}
