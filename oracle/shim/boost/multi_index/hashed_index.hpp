// oracle shim: forwards to the single stand-in header (test infrastructure)
#pragma once
#include <boost/multi_index_container.hpp>
