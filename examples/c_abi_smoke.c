/* c_abi_smoke.c -- the drop-in boundary from plain C (C99, no CUDA headers, no C++): what a P/Invoke or cgo binding sees.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -I include examples/c_abi_smoke.c -o c_abi_smoke \
 *       -L ppo-bipedalwalker_b200/lib -lwalker_b200 -Wl,-rpath,$PWD/ppo-bipedalwalker_b200/lib
 *
 * One Environment.Update-equivalent (Environment.cs:64-92) for 4 walkers, one per floor material, through wb_env_step with
 * ordinary malloc'ed buffers, then again with wb_host_pin'ed buffers (zero-copy path): both must print the same observations.
 * Exit status: 0 = ran on a B200; 3 = no sm_100 device (the library refused loudly, as designed: there is no CPU fallback);
 * 1 = anything else. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "walker_b200.h"

static int fail(const char* what, int32_t rc) {
  char msg[512];
  wb_last_error(msg, sizeof msg);
  fprintf(stderr, "%s: status %d: %s\n", what, (int)rc, msg);
  return rc == WB_ERR_NO_DEVICE ? 3 : 1;
}

int main(void) {
  enum { N = 4 };
  int32_t rc = wb_init(0);
  if (rc != WB_OK) return fail("wb_init", rc);

  wb_hyperparams hp;
  if ((rc = wb_hyperparams_default(&hp)) != WB_OK) return fail("wb_hyperparams_default", rc);
  const uint8_t floors[N] = {WB_ICE, WB_WOOD, WB_RUBBER, WB_METAL};
  float actions[N * WB_ACT];
  for (int i = 0; i < N * WB_ACT; i++) actions[i] = (float)((i * 7) % 11) / 5.0f - 1.0f; /* in [-1, 1] */

  float first_obs[N * WB_OBS];
  for (int pass = 0; pass < 2; pass++) {
    wb_env_batch* env = NULL;
    if ((rc = wb_env_create(N, floors, NULL, &hp, &env)) != WB_OK) return fail("wb_env_create", rc);
    float* a = (float*)malloc(sizeof actions);
    float* obs = (float*)malloc(sizeof(float) * N * WB_OBS);
    float* reward = (float*)malloc(sizeof(float) * N);
    uint8_t* done = (uint8_t*)malloc(N);
    if (!a || !obs || !reward || !done) return 1;
    memcpy(a, actions, sizeof actions);
    if (pass == 1) { /* page-lock the caller's buffers: wb_env_step then reads / writes them directly from the kernel */
      if ((rc = wb_host_pin(a, sizeof actions)) != WB_OK) return fail("wb_host_pin", rc);
      if ((rc = wb_host_pin(obs, sizeof(float) * N * WB_OBS)) != WB_OK) return fail("wb_host_pin", rc);
      if ((rc = wb_host_pin(reward, sizeof(float) * N)) != WB_OK) return fail("wb_host_pin", rc);
      if ((rc = wb_host_pin(done, N)) != WB_OK) return fail("wb_host_pin", rc);
    }
    for (int t = 0; t < 3; t++)
      if ((rc = wb_env_step(env, a, 1.0f / 60.0f, 1, obs, reward, done)) != WB_OK) return fail("wb_env_step", rc);
    if (pass == 0) {
      memcpy(first_obs, obs, sizeof first_obs);
      for (int e = 0; e < N; e++)
        printf("walker %d: hip (%.4f, %.4f) reward %.5f done %d\n", e, obs[e * WB_OBS], obs[e * WB_OBS + 1], reward[e], (int)done[e]);
    } else {
      if (memcmp(first_obs, obs, sizeof first_obs) != 0) {
        fprintf(stderr, "zero-copy and staged observations differ\n");
        return 1;
      }
      if ((rc = wb_host_unpin(a)) != WB_OK || (rc = wb_host_unpin(obs)) != WB_OK || (rc = wb_host_unpin(reward)) != WB_OK ||
          (rc = wb_host_unpin(done)) != WB_OK)
        return fail("wb_host_unpin", rc);
    }
    free(a);
    free(obs);
    free(reward);
    free(done);
    if ((rc = wb_env_destroy(env)) != WB_OK) return fail("wb_env_destroy", rc);
  }
  printf("c_abi_smoke: staged and zero-copy paths agree (%s)\n", wb_version());
  return 0;
}
