// Test-infrastructure stub: stands in for a deal.II header that the reference's
// src/equation.h includes but whose contents the flux arithmetic never touches.
#include "../../dealii_stub_core.h"
