#include "../thread.hpp"
