/* declaration-only stub: lets the reference radio.c compile for the oracle; nothing here is ever called */
#ifndef STUB_OPUS_H
#define STUB_OPUS_H
#include <stdint.h>
typedef struct OpusEncoder OpusEncoder;
typedef int32_t opus_int32;
typedef int16_t opus_int16;
#endif
