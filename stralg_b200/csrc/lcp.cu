// lcp.cu -- placeholder, replaced below
#include "engine.h"
namespace b200sa {
void build_lcp(DeviceIndex &ix) { (void)ix; throw std::runtime_error("LCP not implemented yet"); }
}
