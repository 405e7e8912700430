// TEST INFRASTRUCTURE ONLY (oracle/ref_shim): declarations live in ../../libint2.hpp
#pragma once
#include <libint2.hpp>
