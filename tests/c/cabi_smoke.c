/* Plain-C consumer of include/probpose_b200.h: proves that the boundary is usable without C++ or Python.
 * Built and run by tests/test_cabi_exports.py (no GPU needed: only argument checking is exercised). */
#include <stdio.h>
#include <string.h>

#include "probpose_b200.h"

int main(void) {
  pp_encode_params ep;
  pp_decode_params dp;
  pp_loss_params lp;
  int rc;
  memset(&ep, 0, sizeof ep);
  memset(&dp, 0, sizeof dp);
  memset(&lp, 0, sizeof lp);
  if (pp_version() <= 0) return 1;
  rc = pp_encode(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
  if (rc == PP_OK || strlen(pp_last_error_string()) == 0) return 2;
  lp.B = 1; lp.K = 1; lp.H = 4; lp.W = 4; lp.dtype = 7;
  rc = pp_oks_loss_forward(&lp, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 1.0f, NULL, NULL, 0, NULL, NULL);
  if (rc == PP_OK) return 3;
  rc = pp_sparsemax_tail(NULL, NULL, NULL, PP_F32, 1, 0, 0.5f, 1.0f, NULL);
  if (rc == PP_OK) return 4;
  {
    pp_mailbox mb;
    memset(&mb, 0, sizeof mb);
    if (pp_mailbox_block_bytes(85) != 85 * 56 + 8 + 16) return 5;      /* records, pad to 16, loss + flag + pad */
    if (pp_mailbox_block_bytes(4352) % 16 != 0) return 6;
    if (pp_mailbox_bytes(85, 8, 4) != 4 * 8 * pp_mailbox_block_bytes(85) + 4 * 8 * 4) return 6;   /* blocks + acknowledgements */
    if (pp_mailbox_state_words(4) != 21 || pp_mailbox_ack(&mb, 1u, NULL) == PP_OK) return 6;
    rc = pp_pack_records(4, NULL, NULL, NULL, NULL, NULL, NULL, 1.0f, NULL, NULL, NULL);
    if (rc == PP_OK) return 7;
    rc = pp_mailbox_commit(&mb, 4, NULL, NULL);                          /* inconsistent (all-zero) mailbox */
    if (rc == PP_OK) return 8;
    if (pp_decode_expected_last_kernel() != -1) return 9;                /* nothing decoded on this thread yet */
    printf("sizeof mailbox: %zu\n", sizeof mb);
  }
  printf("sizeof encode/decode/loss params: %zu %zu %zu; version %d; last error: %s\n", sizeof ep, sizeof dp, sizeof lp,
         pp_version(), pp_last_error_string());
  return 0;
}
