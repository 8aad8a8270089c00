// icprep.cu -- internal-constraint projection products (ihwbc.py:124-146) -- placeholder
#include "pnc_launch.h"
namespace pnc {
void launch_ic_prepare(const DevProblem*, const DevProblem&, const WsPtrs&, const FrameMap&, int, int, int, cudaStream_t) {}
}
