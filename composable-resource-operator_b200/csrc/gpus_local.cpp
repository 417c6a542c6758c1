// gpus_local.cpp — the Kube / Exec seams of gpus.hpp answered ON the node: what a node agent that links
// libcroprobe does instead of the operator's SPDY execs into other pods (internal/utils/gpus.go:788-815).
#include <poll.h>
#include <signal.h>
#include <time.h>
#include <errno.h>
#include <spawn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

#include "detach.hpp"
#include "gpus.hpp"
#include "identity.hpp"
#include "nvml_ops.hpp"

extern char** environ;

namespace cro {
namespace gpus {

namespace {

// The dry-run gate is an ALLOW-list: only the argv shapes known to read (the queries and status commands of
// internal/utils/gpus.go: --query-gpu :886, --query-compute-apps :131, `drain -p <bus> -q` :968, lsmod :1105) are
// executed while allow_mutation is false; anything else — including a command added later that nobody classified —
// is logged as skipped.
bool read_only(const std::vector<std::string>& argv) {
    if (argv.empty()) return false;
    const std::string& exe = argv[0];
    auto base_is = [&](const char* name) {
        const size_t slash = exe.rfind('/');
        return (slash == std::string::npos ? exe : exe.substr(slash + 1)) == name;
    };
    if (base_is("lsmod")) return argv.size() == 1;
    if (base_is("nvidia-smi")) {
        if (argv.size() == 3 && argv[2] == "--format=csv,noheader,nounits" &&
            (argv[1].compare(0, 12, "--query-gpu=") == 0 || argv[1].compare(0, 21, "--query-compute-apps=") == 0))
            return true;
        if (argv.size() == 5 && argv[1] == "drain" && argv[2] == "-p" && argv[4] == "-q") return true;
    }
    return false;
}

// nvidia-smi's answer shaped as an exec result: stdout text, and a non-zero exit as kubectl-exec words it.
bool answer_with_nvml(const std::vector<std::string>& argv, const std::string& lib, ExecResult* r) {
    if (argv.empty()) return false;
    const size_t slash = argv[0].rfind('/');
    if ((slash == std::string::npos ? argv[0] : argv[0].substr(slash + 1)) != "nvidia-smi") return false;
    nvml::Reply rep;
    if (argv.size() == 3 && argv[1] == "--query-compute-apps=gpu_uuid,process_name" && argv[2] == "--format=csv,noheader,nounits")
        rep = nvml::ComputeApps(lib);
    else if (argv.size() == 3 && argv[1].compare(0, 12, "--query-gpu=") == 0 && argv[2] == "--format=csv,noheader,nounits")
        rep = nvml::QueryGpu(lib, argv[1].substr(12));
    else if (argv.size() == 5 && argv[1] == "drain" && argv[2] == "-p" && argv[4] == "-q")
        rep = nvml::DrainQuery(lib, argv[3]);
    else if (argv.size() == 5 && argv[1] == "drain" && argv[2] == "-p" && argv[4] == "-r")
        rep = nvml::DrainRemove(lib, argv[3]);
    else if (argv.size() == 6 && argv[1] == "drain" && argv[2] == "-p" && argv[4] == "-m" && (argv[5] == "0" || argv[5] == "1"))
        rep = nvml::DrainModify(lib, argv[3], argv[5] == "1");
    else if (argv.size() == 5 && argv[1] == "-i" && argv[3] == "-pm" && (argv[4] == "0" || argv[4] == "1"))
        rep = nvml::SetPersistence(lib, argv[2], argv[4] == "1");
    else
        return false;
    if (!rep.available) return false;
    r->std_out = rep.std_out;
    if (rep.exit_code != 0) {
        r->failed = true;
        r->exec_err = "command terminated with exit code " + std::to_string(rep.exit_code);
    }
    return true;
}

long long now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

// posix_spawn + pipes; the error text is kubectl-exec's ("command terminated with exit code N").  The child has
// `deadline_ms` to finish (the reference's exec is bound to the reconcile's context; a wedged nvidia-smi on a GPU that
// is mid-drain must not hang the agent thread): on expiry it is killed, reaped, and the call fails like a cancelled
// context does ("context deadline exceeded").
ExecResult spawn(const std::vector<std::string>& argv, int deadline_ms) {
    ExecResult r;
    if (argv.empty()) { r.failed = true; r.exec_err = "empty command"; return r; }
    int out[2], err[2];
    if (pipe(out) != 0) { r.failed = true; r.exec_err = "pipe failed"; return r; }
    if (pipe(err) != 0) { close(out[0]); close(out[1]); r.failed = true; r.exec_err = "pipe failed"; return r; }
    posix_spawn_file_actions_t fa;
    posix_spawn_file_actions_init(&fa);
    posix_spawn_file_actions_adddup2(&fa, out[1], 1);
    posix_spawn_file_actions_adddup2(&fa, err[1], 2);
    posix_spawn_file_actions_addclose(&fa, out[0]);
    posix_spawn_file_actions_addclose(&fa, err[0]);
    std::vector<char*> av;
    for (const auto& a : argv) av.push_back(const_cast<char*>(a.c_str()));
    av.push_back(nullptr);
    pid_t pid = 0;
    const int rc = posix_spawn(&pid, argv[0].c_str(), &fa, nullptr, av.data(), environ);
    posix_spawn_file_actions_destroy(&fa);
    close(out[1]);
    close(err[1]);
    if (rc != 0) {
        close(out[0]);
        close(err[0]);
        r.failed = true;
        r.exec_err = std::string("exec: \"") + argv[0] + "\": " + strerror(rc);
        return r;
    }
    pollfd fds[2] = {{out[0], POLLIN, 0}, {err[0], POLLIN, 0}};
    int open_fds = 2;
    char buf[4096];
    const long long until = now_ms() + (deadline_ms > 0 ? deadline_ms : 60000);
    bool timed_out = false;
    while (open_fds > 0) {
        const long long left = until - now_ms();
        if (left <= 0) { timed_out = true; break; }
        const int pr = poll(fds, 2, (int)std::min<long long>(left, 1000));
        if (pr < 0 && errno == EINTR) continue;
        if (pr < 0) break;
        if (pr == 0) continue;
        for (int i = 0; i < 2; ++i) {
            if (fds[i].fd < 0 || !(fds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            const ssize_t n = read(fds[i].fd, buf, sizeof buf);
            if (n > 0) (i == 0 ? r.std_out : r.std_err).append(buf, (size_t)n);
            else { close(fds[i].fd); fds[i].fd = -1; --open_fds; }
        }
    }
    for (pollfd& f : fds)
        if (f.fd >= 0) close(f.fd);               // poll() failed mid-way: do not leak the read ends
    int status = 0;
    if (!timed_out) {
        // the pipes are closed; the process itself gets the rest of the deadline to exit
        for (;;) {
            const pid_t w = waitpid(pid, &status, WNOHANG);
            if (w == pid || (w < 0 && errno != EINTR)) break;
            if (now_ms() > until) { timed_out = true; break; }
            usleep(500);
        }
    }
    if (timed_out) {
        kill(pid, SIGKILL);
        waitpid(pid, &status, 0);
        r.failed = true;
        r.exec_err = "context deadline exceeded";
        return r;
    }
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        r.failed = true;
        r.exec_err = "command terminated with exit code " + std::to_string(WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status));
    }
    return r;
}

}  // namespace

LocalExec::LocalExec(const Options& o) : o_(o) {}

void LocalExec::Sleep(int s) {
    if (o_.allow_mutation) sleep((unsigned)s);
}

ExecResult LocalExec::Run(const Pod&, const std::string&, const ExecRequest& req) {
    LogEntry le;
    le.kind = (int)req.kind;
    le.argv = req.kind == ExecRequest::Command ? req.argv : ScanAsCommand(req);
    ExecResult r;
    switch (req.kind) {
        case ExecRequest::FdScan:
            le.how = "native";
            r.std_out = detach::ScanDeviceFileHolders(o_.proc_root, req.target, req.rke2_format);
            break;
        case ExecRequest::CmdlineScan:
            le.how = "native";
            r.std_out = detach::ScanCmdlineFor(o_.proc_root, req.target);
            break;
        case ExecRequest::ProcScan: {
            le.how = "native";
            for (const identity::ProcGpu& g : identity::ScanProc(o_.proc_root.empty() ? std::string("/proc") : o_.proc_root))
                r.std_out += g.minor + "," + g.uuid + "," + g.bus + "\n";
            break;
        }
        case ExecRequest::Command: {
            std::vector<std::string> argv = req.argv;
            if (argv.size() > 2 && argv[0] == "/bin/chroot" && argv[1] == "/host-root") argv.erase(argv.begin(), argv.begin() + 2);
            // `nvidia-smi --query-gpu=<fields> --format=csv,noheader,nounits` from the devices the probe context already
            // enumerated (NVML + /proc): no process spawn — and it knows device_minor, which driver 580's nvidia-smi
            // refuses to print ("not a valid field to query"), so the reference's own DrainGPU cannot even start there
            if (o_.devs && o_.n_devs >= 0 && argv.size() == 3 && argv[1].compare(0, 12, "--query-gpu=") == 0) {
                std::string out, err;
                if (identity::EmitCsv(o_.devs, o_.n_devs, argv[1].substr(12), &out, &err) == 0) {
                    le.how = "native";
                    r.std_out = out;
                    break;
                }
            }
            if (!read_only(argv) && !o_.allow_mutation) {
                le.how = "skipped (dry run)";
                break;
            }
            // the detach side's nvidia-smi invocations through this process's NVML session (nvml_ops.hpp); without
            // NVML here (or with native_nvml off) the command is spawned like any other
            if (o_.native_nvml && answer_with_nvml(argv, o_.nvml_lib, &r)) {
                le.how = "native";
                break;
            }
            le.how = "spawned";
            r = spawn(argv, o_.exec_deadline_ms);
            break;
        }
    }
    le.failed = r.failed;
    log.push_back(le);
    return r;
}

Error LocalKube::GetClusterPolicy(bool* found, bool* set, bool* enabled) {
    *found = driver_container_;       // a containerised driver (gpu-operator) or the host's own (RKE2 flavour)
    *set = driver_container_;
    *enabled = driver_container_;
    return Error::Nil();
}

Error LocalKube::ListPods(std::vector<Pod>* out) {
    // one stand-in per role: on the node every "pod" is this process
    out->push_back({"local", "nvidia-driver-daemonset-local", node_, {{"app.kubernetes.io/component", "nvidia-driver"}}, {"local"}});
    out->push_back({"local", "nvidia-dra-driver-gpu-kubelet-plugin-local", node_, {{"app.kubernetes.io/name", "nvidia-dra-driver-gpu"}}, {"local"}});
    out->push_back({"local", "cro-node-agent-local", node_, {{"app", "cro-node-agent"}}, {"local"}});
    return Error::Nil();
}

}  // namespace gpus
}  // namespace cro
