/* A plain C99 host of the C-ABI (what a compiled reference-side caller would look like): include/b200rec.h compiles as
 * C, the library links with -lb200rec, and argument errors come back as return codes + b200rec_last_error() before any
 * CUDA call is made, so this runs without a GPU.  Built and run by tests/test_cabi_and_host.py. */
#include <stdio.h>
#include <string.h>

#include "b200rec.h"

int main(void) {
  size_t ws;
  if (b200rec_version() != 1) return 2;
  ws = b200rec_topk_workspace_bytes(0, 128, 4, 10);              /* empty catalogue */
  if (ws != 0 || strstr(b200rec_last_error(), "empty") == NULL) return 3;
  ws = b200rec_topk_workspace_bytes(10000000, 128, 4096, 100);   /* BASELINE config 3 */
  if (ws == 0) return 4;
  if (b200rec_flat_ip_topk(NULL, 0, 0, NULL, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) == 0) return 5;
  if (strstr(b200rec_last_error(), "null") == NULL) return 6;
  if (b200rec_launch_count() != 0) return 7;                     /* nothing was launched */
  printf("config-3 workspace: %zu bytes\n", ws);
  return 0;
}
