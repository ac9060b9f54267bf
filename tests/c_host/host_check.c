/* A plain-C host of the C ABI (include/lsnf.h), the binding a non-Python maintainer would write: dlopen the
 * library, resolve the entry points by name, build a plan for a BASELINE configuration and walk its stage
 * tables.  Built and run by tests/test_c_host.py (gcc, C99); prints `key value` lines that the test compares
 * with the ctypes binding.  No GPU is needed for anything but the last call, which must FAIL LOUDLY without
 * one (there is no CPU fallback behind the ABI). */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "lsnf.h"

#define RESOLVE(name)                                          \
  *(void**)(&p_##name) = dlsym(lib, #name);                    \
  if (!p_##name) {                                             \
    fprintf(stderr, "missing symbol %s\n", #name);             \
    return 2;                                                  \
  }

int main(int argc, char** argv) {
  int (*p_lsnf_abi_version)(void);
  const char* (*p_lsnf_last_error)(void);
  int (*p_lsnf_plan_create)(const lsnf_config*, lsnf_plan**);
  void (*p_lsnf_plan_destroy)(lsnf_plan*);
  size_t (*p_lsnf_workspace_bytes)(const lsnf_plan*);
  int (*p_lsnf_plan_num_stages)(const lsnf_plan*);
  int (*p_lsnf_plan_stage_info)(const lsnf_plan*, int32_t, lsnf_stage_info*);
  int (*p_lsnf_plan_stage_launch_info)(const lsnf_plan*, int32_t, int32_t, lsnf_launch_info*);
  int (*p_lsnf_langevin_launch_count)(const lsnf_plan*, int32_t);
  int (*p_lsnf_langevin_run)(lsnf_plan*, const float*, const float*, int32_t, float, float, int32_t, const float*,
                             uint64_t, uint64_t, float*, float*, lsnf_stream);
  void* lib;
  lsnf_config cfg;
  lsnf_plan* plan = NULL;
  lsnf_stage_info si;
  lsnf_launch_info li;
  int i, n, rc;

  if (argc < 2) return 1;
  lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 2;
  }
  RESOLVE(lsnf_abi_version)
  RESOLVE(lsnf_last_error)
  RESOLVE(lsnf_plan_create)
  RESOLVE(lsnf_plan_destroy)
  RESOLVE(lsnf_workspace_bytes)
  RESOLVE(lsnf_plan_num_stages)
  RESOLVE(lsnf_plan_stage_info)
  RESOLVE(lsnf_plan_stage_launch_info)
  RESOLVE(lsnf_langevin_launch_count)
  RESOLVE(lsnf_langevin_run)

  printf("abi_version %d\n", p_lsnf_abi_version());
  printf("header_abi_version %d\n", LSNF_ABI_VERSION);
  printf("sizeof_config %zu\n", sizeof(lsnf_config));
  printf("sizeof_tap %zu\n", sizeof(lsnf_tap));
  printf("sizeof_stage_info %zu\n", sizeof(lsnf_stage_info));
  printf("sizeof_launch_info %zu\n", sizeof(lsnf_launch_info));
  printf("offsetof_config_leak %zu\n", offsetof(lsnf_config, leak));
  printf("offsetof_config_train %zu\n", offsetof(lsnf_config, train));
  printf("offsetof_stage_taps %zu\n", offsetof(lsnf_stage_info, taps));
  printf("offsetof_stage_passes %zu\n", offsetof(lsnf_stage_info, passes));
  printf("offsetof_stage_a_offset %zu\n", offsetof(lsnf_stage_info, a_offset));
  printf("offsetof_stage_flops %zu\n", offsetof(lsnf_stage_info, flops));
  printf("offsetof_launch_smem_bytes %zu\n", offsetof(lsnf_launch_info, smem_bytes));

  /* an invalid configuration is rejected with a message (nz must be even, model.py:383) */
  memset(&cfg, 0, sizeof cfg);
  cfg.arch = LSNF_ARCH_CIFAR10; cfg.batch = 100; cfg.nz = 127; cfg.ngf = 128; cfg.nc = 3;
  cfg.f_depth = 5; cfg.f_width = 64; cfg.f_permutation = 2; cfg.f_coupling = 1; cfg.leak = 0.2f;
  rc = p_lsnf_plan_create(&cfg, &plan);
  printf("odd_nz_rc %d\n", rc);
  printf("odd_nz_has_message %d\n", (int)(strlen(p_lsnf_last_error()) > 0));

  /* BASELINE config 2 (CIFAR-10): train.py --dataset cifar10 --nz 128 --ngf 128 --f_width 64, batch 100 */
  cfg.nz = 128;
  rc = p_lsnf_plan_create(&cfg, &plan);
  printf("create_rc %d\n", rc);
  if (rc != LSNF_OK) {
    fprintf(stderr, "%s\n", p_lsnf_last_error());
    return 3;
  }
  printf("workspace_bytes %zu\n", p_lsnf_workspace_bytes(plan));
  n = p_lsnf_plan_num_stages(plan);
  printf("num_stages %d\n", n);
  printf("launches_per_40_steps %d\n", p_lsnf_langevin_launch_count(plan, 40));
  for (i = 0; i < n; ++i) {
    if (p_lsnf_plan_stage_info(plan, i, &si) != LSNF_OK) return 4;
    if (p_lsnf_plan_stage_launch_info(plan, i, 148, &li) != LSNF_OK) return 5;
    printf("stage %d kind %d layer %d block_n %d passes %d k_splits %d flops %lld kernel %d grid %d %d %d smem %d "
           "tmem %d\n", i, si.kind, si.layer, si.block_n, si.passes, si.k_splits, (long long)si.flops, li.kernel,
           li.grid_x, li.grid_y, li.grid_z, li.smem_bytes, li.tmem_columns);
  }
  /* compute before lsnf_plan_bind / without a device: must fail with a status and a message, never "succeed" on
   * the host */
  rc = p_lsnf_langevin_run(plan, NULL, NULL, 40, 0.1f, 0.3f, 1, NULL, 1u, 0u, NULL, NULL, NULL);
  printf("unbound_run_rc %d\n", rc);
  printf("unbound_run_has_message %d\n", (int)(strlen(p_lsnf_last_error()) > 0));
  p_lsnf_plan_destroy(plan);
  dlclose(lib);
  return 0;
}
