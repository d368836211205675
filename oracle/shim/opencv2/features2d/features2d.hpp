// stand-in header (test infrastructure): see ../opencv_shim.hpp
#include "../opencv_shim.hpp"
