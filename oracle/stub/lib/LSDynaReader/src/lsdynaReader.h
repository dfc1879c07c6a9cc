/* TEST INFRASTRUCTURE (oracle): forwards the relative include used by
 * src/common/Domain_d.C:44 to the stub in oracle/stub/inc. */
#include "../../../inc/lsdynaReader.h"
