// placeholder — replaced by the supernodal LDLt implementation
#include "fpsb_internal.h"
namespace fpsb {
void ldlt_free(Handle *h) { (void)h; }
}
