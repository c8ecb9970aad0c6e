/* Plain-C consumer of libkfnet_b200.so: includes the public header, resolves every entry point with dlsym and makes
 * the calls that need no GPU.  Built and run by tests/test_cpu_host.py::test_c_consumer (gcc, -ldl). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "kfnet_b200.h"

#define SYM(name)                                             \
  do {                                                        \
    if (!dlsym(lib, #name)) {                                 \
      fprintf(stderr, "missing symbol %s\n", #name);          \
      return 2;                                               \
    }                                                         \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: smoke <libkfnet_b200.so>\n"); return 64; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  SYM(kfn_version); SYM(kfn_last_error); SYM(kfn_create); SYM(kfn_destroy); SYM(kfn_load_weight);
  SYM(kfn_finalize_weights); SYM(kfn_workspace_bytes); SYM(kfn_scoordnet_forward); SYM(kfn_oflow_feat_forward);
  SYM(kfn_diff_volume); SYM(kfn_oflownet_forward); SYM(kfn_kalman_scan); SYM(kfn_flow_warp_kalman);
  SYM(kfn_bilinear_sampler); SYM(kfn_apply_transform); SYM(kfn_kalman_update); SYM(kfn_sequence_extra_bytes);
  SYM(kfn_eval_sequences); SYM(kfn_eval_sequences_host); SYM(kfn_eval_stream_host); SYM(kfn_conv_layer); SYM(kfn_launch_count);
  SYM(kfn_train_begin); SYM(kfn_train_step); SYM(kfn_train_get); SYM(kfn_train_commit); SYM(kfn_train_set_step);

  const char* (*version)(void) = (const char* (*)(void))dlsym(lib, "kfn_version");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(lib, "kfn_last_error");
  int (*create)(kfn_handle**, int, int, int, int) = (int (*)(kfn_handle**, int, int, int, int))dlsym(lib, "kfn_create");
  int (*destroy)(kfn_handle*) = (int (*)(kfn_handle*))dlsym(lib, "kfn_destroy");
  printf("version: %s\n", version());

  kfn_handle* h = NULL;
  int rc = create(&h, 0, 99, 480, 640);                 /* unknown mode: argument error before any CUDA call */
  if (rc != KFN_ERR_ARG || strlen(last_error()) == 0) { fprintf(stderr, "bad mode not rejected (rc %d)\n", rc); return 3; }
  rc = create(&h, 0, KFN_MODE_BF16, 481, 640);          /* image size must be a multiple of 8 */
  if (rc != KFN_ERR_ARG) { fprintf(stderr, "bad size not rejected (rc %d)\n", rc); return 4; }
  rc = create(&h, 0, KFN_MODE_BF16, 480, 640);
  if (rc == KFN_OK) {                                   /* a B200 is present */
    printf("create: ok\n");
    if (destroy(h) != KFN_OK) return 5;
  } else {                                              /* no GPU / not sm_100: must fail loudly, never fall back */
    printf("create: rc %d (%s)\n", rc, last_error());
    if (rc != KFN_ERR_CUDA && rc != KFN_ERR_ARG) return 6;
    if (strlen(last_error()) == 0) return 7;
  }
  if (destroy(NULL) != KFN_OK) return 8;                /* destroying a null handle is a no-op */
  dlclose(lib);
  printf("c-abi smoke: ok\n");
  return 0;
}
