// detach.cpp — see detach.hpp.
#include "detach.hpp"

#include <iterator>

#include <dirent.h>
#include <limits.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

#include "identity.hpp"

namespace cro {
namespace detach {

using identity::Split;
using identity::ToLower;
using identity::TrimSpace;

static std::string errText(const char* exec_err) { return exec_err ? exec_err : "<nil>"; }

Error CheckNoGPULoadsFromOutput(const std::string& stdOut, const std::string& stdErr, const char* exec_err,
                                const std::string& podName, const std::string& targetNodeName,
                                const std::string* targetGPUUUID, bool driverEnabled) {
    if (TrimSpace(stdOut) == "No devices were found") return Error::Nil();   // gpus.go:146
    if (!stdErr.empty() || exec_err != nullptr)                              // :150
        return Error::New("run nvidia-smi in pod '" + podName + "' to check gpu loads failed: '" + errText(exec_err) +
                          "', stderr: '" + stdErr + "', stdout: '" + stdOut + "'");
    struct App { std::string uuid, name; };
    std::vector<App> apps;
    for (const std::string& line : Split(TrimSpace(stdOut), "\n")) {         // :156-170
        if (line.empty()) continue;
        const std::vector<std::string> parts = Split(line, ",");
        if (parts.size() < 2)   // parts[1] is indexed unguarded: Go panics
            return Error::New("runtime error: index out of range [1] with length " + std::to_string(parts.size()));
        apps.push_back({TrimSpace(parts[0]), TrimSpace(parts[1])});
    }
    // fmt %v of []AccountedAppInfo with its String() method (gpus.go:50-52)
    std::string listed = "[";
    for (size_t i = 0; i < apps.size(); ++i) {
        if (i) listed += " ";
        listed += "GPUUUID: '" + apps[i].uuid + "', ProcessName: '" + apps[i].name + "'";
    }
    listed += "]";
    if (!driverEnabled) {                                                     // :172-178 (RKE2)
        if (!targetGPUUUID) return Error::New("runtime error: invalid memory address or nil pointer dereference");
        for (const App& a : apps)
            if (a.uuid == *targetGPUUUID)
                return Error::New("found gpu load on gpu '" + *targetGPUUUID + "': " + listed);
        return Error::Nil();
    }
    if (!apps.empty())                                                        // :179-182
        return Error::New("found gpu loads on node '" + targetNodeName + "': '" + listed + "'");
    return Error::Nil();
}

Error checkGPUDrainStatusFromOutput(const std::string& stdOut, const std::string& stdErr, const char* exec_err,
                                    const std::string& targetNodeName, const std::string& targetGPUBusID,
                                    bool* draining) {
    *draining = false;
    const std::string busID = TrimSpace(targetGPUBusID);
    if (busID.empty()) return Error::New("target GPU bus ID is empty");                        // :966-968
    if (exec_err != nullptr || !stdErr.empty())                                                // :980-982
        return Error::New("check gpu drain status command failed: '" + errText(exec_err) + "', stderr: '" + stdErr +
                          "', stdout: '" + stdOut + "'");
    const std::string trimmed = TrimSpace(stdOut);
    if (trimmed.empty())
        return Error::New("nvidia-smi drain query returned empty output (node=" + targetNodeName + ", busID=" + busID + ")");
    for (const std::string& line : Split(trimmed, "\n")) {                                     // :989-1007
        const std::string lower = ToLower(TrimSpace(line));
        if (lower.find("drain") == std::string::npos) continue;
        const size_t idx = lower.find(':');
        if (idx == std::string::npos) continue;
        std::string status = TrimSpace(lower.substr(idx + 1));
        size_t b = 0, e = status.size();                     // strings.Trim(status, ".")
        while (b < e && status[b] == '.') ++b;
        while (e > b && status[e - 1] == '.') --e;
        status = status.substr(b, e - b);
        if (status.find("not draining") != std::string::npos) return Error::Nil();
        if (status.find("draining") != std::string::npos) {
            *draining = true;
            return Error::Nil();
        }
    }
    return Error::New("nvidia-smi drain query did not contain recognizable drain state (node=" + targetNodeName +
                      ", busID=" + busID + ", raw=" + trimmed + ")");
}

Error CheckDeviceFileScanResult(const std::string& stdOut, const std::string& stdErr, const char* exec_err, bool rke2) {
    if (rke2) {   // gpus.go:286-291
        if (exec_err != nullptr || !stdErr.empty())
            return Error::New("deatch command 'check /dev/nvidiaX' failed: '" + errText(exec_err) + "', stderr: '" + stdErr +
                              "', stdout: '" + stdOut + "'");
        if (!stdOut.empty())
            return Error::New("check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: " + stdOut);
        return Error::Nil();
    }
    if (!stdErr.empty() || exec_err != nullptr)   // :468-470
        return Error::New("check /dev/nvidiaX command failed: '" + errText(exec_err) + "', stderr: '" + stdErr + "'");
    if (!stdOut.empty())                          // :471-473
        return Error::New("check /dev/nvidiaX command failed: there is a process " + stdOut + " occupied the nvidiaX file");
    return Error::Nil();
}

std::string ScanDeviceFileHolders(const std::string& proc_root, const std::string& target, bool rke2) {
    const std::string root = proc_root.empty() ? "/proc" : proc_root;
    std::vector<std::string> pids;
    if (DIR* d = opendir(root.c_str())) {
        while (dirent* e = readdir(d))
            if (e->d_name[0] >= '0' && e->d_name[0] <= '9') pids.push_back(e->d_name);
        closedir(d);
    }
    std::sort(pids.begin(), pids.end());   // the shell glob /proc/[0-9]* expands in lexical order
    struct stat tst;
    const bool have_target = stat(target.c_str(), &tst) == 0;
    std::string out;
    bool first = true;
    for (const std::string& pid : pids) {
        const std::string dir = root + "/" + pid;
        std::string comm = "[unknown]";
        {
            std::ifstream f(dir + "/comm");
            std::string line;
            if (f && std::getline(f, line)) comm = line;
            else if (rke2) continue;   // `CMD_NAME=$(cat ...) || continue`
        }
        std::vector<std::string> fds;
        if (DIR* fd = opendir((dir + "/fd").c_str())) {
            while (dirent* e = readdir(fd))
                if (e->d_name[0] != '.') fds.push_back(e->d_name);
            closedir(fd);
        }
        std::sort(fds.begin(), fds.end());
        for (const std::string& n : fds) {
            const std::string link = dir + "/fd/" + n;
            if (rke2) {   // find -L ... -samefile "$TARGET_FILE"
                struct stat st;
                if (!have_target || stat(link.c_str(), &st) != 0) continue;
                if (st.st_dev != tst.st_dev || st.st_ino != tst.st_ino) continue;
                if (!first) out += ", ";
                first = false;
                out += pid + " " + comm;
            } else {      // readlink -f "$FD_SYMLINK" == "$TARGET_FILE" -> echo comm; exit 0
                struct stat lst;
                if (lstat(link.c_str(), &lst) != 0 || !S_ISLNK(lst.st_mode)) continue;
                char buf[PATH_MAX];
                if (!realpath(link.c_str(), buf)) continue;
                if (target == buf) return comm + "\n";
            }
        }
    }
    return out;
}

std::string ScanCmdlineFor(const std::string& proc_root, const std::string& target) {
    const std::string root = proc_root.empty() ? "/proc" : proc_root;
    const std::string self = std::to_string((long long)getpid()), parent = std::to_string((long long)getppid());
    std::vector<std::string> pids;
    if (DIR* d = opendir(root.c_str())) {
        while (dirent* e = readdir(d))
            if (e->d_name[0] >= '0' && e->d_name[0] <= '9') pids.push_back(e->d_name);
        closedir(d);
    }
    for (const std::string& pid : pids) {
        if (proc_root.empty() && (pid == self || pid == parent)) continue;   // the scanner never reports itself
        std::ifstream f(root + "/" + pid + "/cmdline", std::ios::binary);
        if (!f) continue;
        std::string cmd((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        for (char& c : cmd)
            if (c == '\0') c = ' ';                                          // tr '\0' ' '
        if (!cmd.empty() && cmd.find(target) != std::string::npos) return "true\n";
    }
    return "";
}

}  // namespace detach
}  // namespace cro
