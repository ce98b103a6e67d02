/* TEST INFRASTRUCTURE ONLY: empty deal.II stub header (see dealii_stub_core.h) */
#pragma once
