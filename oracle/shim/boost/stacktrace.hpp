#pragma once
#include <ostream>
namespace boost { namespace stacktrace { struct stacktrace {}; inline std::ostream& operator<<(std::ostream& o, const stacktrace&) { return o << "(stacktrace unavailable)"; } } }
