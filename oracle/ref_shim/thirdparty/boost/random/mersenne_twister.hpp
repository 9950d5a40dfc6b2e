// Stand-in (see ../random.hpp) -- TEST INFRASTRUCTURE for oracle/_ref.
#include "boost/random.hpp"
