// compat stand-in: Grid3d.h includes <tf/tf.h> but the hot-path TUs use nothing from it.
// Deliberately empty (and deliberately free of <math.h>, see compat/README.md).
#pragma once
