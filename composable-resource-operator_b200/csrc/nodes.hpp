// nodes.hpp — the DaemonSet restart rule the attach / detach steps call after a
// device appears or disappears (internal/utils/nodes.go:35-76), as a decision
// over a DaemonSet view and a clock; the Get / Update stay with the host.
#pragma once

#include <string>

#include "reconcile.hpp"

namespace cro {
namespace nodes {

using controller::Error;

struct DaemonSetView {                 // appsv1.DaemonSetStatus + the pod-template annotation
    long long DesiredNumberScheduled = 0, NumberReady = 0, CurrentNumberScheduled = 0;
    long long NumberUnavailable = 0, NumberMisscheduled = 0;
    bool hasRestartedAt = false;       // Spec.Template.Annotations["kubectl.kubernetes.io/restartedAt"]
    std::string restartedAt;
};

// time.Parse(time.RFC3339, value): true + Unix seconds, or false + the text of the
// *time.ParseError (`parsing time "error" as "2006-01-02T15:04:05Z07:00": cannot
// parse "error" as "2006"`, pinned by composableresource_controller_test.go:2863).
bool ParseRFC3339(const std::string& value, long long* unixSeconds, long long* nanos, std::string* err);

// time.Now().Format(time.RFC3339) in UTC for a Unix time.
std::string FormatRFC3339UTC(long long unixSeconds);

enum class Restart { Skipped, Restarted };
// RestartDaemonset after the Get succeeded.  *out says whether the reference would
// issue the Update (stamping restartedAt = now); nowUnix/nowNanos are time.Now().
Error RestartDaemonsetDecision(const std::string& ns, const std::string& name, const DaemonSetView& ds,
                               long long nowUnix, long long nowNanos, Restart* out);

}  // namespace nodes
}  // namespace cro
