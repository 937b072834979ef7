/*
 * gabby_b200_host.h -- C view of the host C++ layer (gabby_b200/host/), for callers that cannot
 * include C++ headers (tests, bench.py). The C++ interface itself mirrors gabby's:
 * gabby::inference::{Message, Request, Generator, Llama3Generator::Load}
 * (/root/reference/src/inference/generator.h:16-47) -- see gabby_b200/host/generator.h.
 *
 * All functions return 0 on success; gb_last_error() holds the message otherwise.
 */
#ifndef GABBY_B200_HOST_H_
#define GABBY_B200_HOST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* gb_last_error(void);

/* RoPE (cos, sin) table [max_pos][head_dim/2][2] fp32 for b2l_create, from config.json's
 * rope_theta / rope_scaling (llama3 != 0: apply the llama3 band rescaling). */
int gb_rope_table(double rope_theta, int llama3, double factor, double low_freq_factor, double high_freq_factor,
                  int original_max_position, int head_dim, int max_pos, float* out);

#ifdef __cplusplus
}
#endif
#endif
