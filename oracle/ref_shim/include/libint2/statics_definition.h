// TEST INFRASTRUCTURE ONLY (oracle/ref_shim): nothing to define; see ../libint2.hpp
#pragma once
