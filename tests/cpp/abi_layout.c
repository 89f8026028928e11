/* Prints the layout of the structs the C-ABI shares with its callers, as seen by a C compiler that includes the public
 * header; tests/test_capi_symbols.py compares it with the ctypes mirrors of rustfst_b200/ffi.py. */
#include <stddef.h>
#include <stdio.h>

#include "rustfst_b200.h"

#define FIELD(S, f) printf("%s.%s %zu\n", #S, #f, offsetof(S, f))

int main(void) {
  printf("B200ComposeStats.sizeof %zu\n", sizeof(B200ComposeStats));
  FIELD(B200ComposeStats, states_expanded); FIELD(B200ComposeStats, arcs_out); FIELD(B200ComposeStats, kernel_launches);
  FIELD(B200ComposeStats, emit_launches); FIELD(B200ComposeStats, ms_expand); FIELD(B200ComposeStats, ms_connect);
  FIELD(B200ComposeStats, ms_emit_kernel); FIELD(B200ComposeStats, ms_h2d); FIELD(B200ComposeStats, ms_d2h);
  FIELD(B200ComposeStats, ms_phase_match); FIELD(B200ComposeStats, ms_phase_resolve);
  printf("B200SsspStats.sizeof %zu\n", sizeof(B200SsspStats));
  FIELD(B200SsspStats, arcs_relaxed); FIELD(B200SsspStats, relax_launches); FIELD(B200SsspStats, path);
  FIELD(B200SsspStats, queue_kind); FIELD(B200SsspStats, ms_device); FIELD(B200SsspStats, ms_relax_kernel);
  FIELD(B200SsspStats, ms_h2d); FIELD(B200SsspStats, ms_queue_plan_host); FIELD(B200SsspStats, ms_order_device);
  FIELD(B200SsspStats, order_on_device); FIELD(B200SsspStats, sweep);
  return 0;
}
