#include <cstdint>
