#include "thread.hpp"
