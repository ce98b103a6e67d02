// Test-infrastructure stub for Trilinos Sacado (only #included, never used by the explicit path).
