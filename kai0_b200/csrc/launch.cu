// Process-wide switch of programmatic dependent launch (launch.h).
#include "launch.h"

namespace pi05 {

int& pdl_state() {
  static int s = -1;
  return s;
}

}  // namespace pi05
