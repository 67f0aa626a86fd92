/* Plain C99 client of include/hnm_b200.h: proves the header is C (not C++) and that the shared library links
 * and answers from outside Python.  Host-only entry points; no GPU needed.  Built and run by
 * tests/test_host_cpu.py::test_header_is_c99_and_library_links_from_c. */
#include <stdio.h>
#include "hnm_b200.h"

int main(void) {
  int32_t plan[5];
  int rc = hnm_score_topk_fused_plan(10719 * 128, 825 * 128, plan);
  if (rc != 0) return 10;
  if (hnm_abi_version() != HNM_ABI_VERSION) return 11;
  if (hnm_score_topk_fused_plan(100, 128, plan) != HNM_E_RANGE) return 12;        /* users_padded not a multiple of 128 */
  if (hnm_score_topk_fused_plan(128, 128, (int32_t*)0) != HNM_E_NULL) return 13;
  if (hnm_score_topk_fused_workspace_bytes(10719 * 128, 825 * 128) <= 0) return 14;
  if (hnm_graph_build_workspace_bytes(0, 0, 0) != 0) return 15;
  printf("%s|%s\n", hnm_strerror(0), hnm_strerror(HNM_E_WORKSPACE));
  return 0;
}
