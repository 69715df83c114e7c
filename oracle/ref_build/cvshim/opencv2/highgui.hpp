/* forwarding header of the test-only cv:: shim (oracle/ref_build/cvshim/cvshim.hpp) */
#include "cvshim.hpp"
