// env.hpp — the CRO_* tuning knobs, validated.
//
// The reference validates every environment value it reads and refuses to
// start on anything else, with one wording
// (internal/controller/composableresource_adapter.go:42-45, :64, :67):
//     the env variable DEVICE_RESOURCE_TYPE has an invalid value: '<v>'
// The probe library follows that convention for its own knobs: each one is
// declared here with its legal range, parsed strictly (decimal digits only, no
// sign, no trailing text) ONCE per cro_probe_init, and an illegal value fails
// the init with the reference's sentence instead of being fed to atoi().
#pragma once
#include <string>

namespace cro {
namespace env {

struct Knob {
    const char* name;
    unsigned lo, hi, dflt;
    unsigned multiple_of;   // 0 = any
    const char* what;
};

// Every knob the library reads.  n_out receives the count.
const Knob* table(int* n_out);

// Strict parse of one value against one knob.  Returns false and fills *err with
// "the env variable <NAME> has an invalid value: '<v>'" on anything illegal.
bool parse(const Knob& k, const char* text, unsigned* out, std::string* err);

// Reads every knob from the process environment (unset or empty = default).  False + *err on the first illegal
// one.  Values land in a process-wide snapshot that get() answers from; cro_probe_init calls this.
bool reload(std::string* err);

// Snapshot value of a knob (its default before the first reload()).  Unknown names abort in debug builds and
// return 0 otherwise: every name used in the code base is in table().
unsigned get(const char* name);

}  // namespace env
}  // namespace cro
