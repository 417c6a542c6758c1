// provider.hpp — the two FTI CdiProvider implementations (fabric manager "FM",
// cluster manager "CM") over an injected transport.
//
// Restates internal/cdi/fti/fm/client.go:100-511 and
// internal/cdi/fti/cm/client.go:107-509: what is asked of the fabric (method,
// path, query, JSON body), how each reply — 200, error status, non-JSON — turns
// into (deviceID, CDIDeviceID, error), and the Node -> Metal3Machine ->
// BareMetalHost annotation walk that names the machine.  The sockets, TLS and
// the credential POST of fti/token.go stay with the host (what its reply means is
// restated below: TokenFromReply): `Transport`, `ObjectStore` and `TokenSource` are the seams, so the same code runs against
// the Go operator's clients and against the scripted fabric
// of the parity tests (the reference's httptest server,
// composableresource_controller_test.go:663-930, restated as data).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "fabric.hpp"
#include "reconcile.hpp"

namespace cro {
namespace fabric {

struct HttpRequest {
    std::string method, path, query, body;   // path is relative to FTI_CDI_ENDPOINT ("fabric_manager/api/v1/...")
};
struct HttpReply {
    int status = 200;
    std::string body;
    std::string transport_error;              // client.Do failed (err.Error()); status/body unused
};
class Transport {
public:
    virtual ~Transport() {}
    virtual HttpReply Do(const HttpRequest& req) = 0;
};

struct K8sObject {
    std::string name;
    bool has_annotations = false;
    std::map<std::string, std::string> annotations;
    std::string provider_id;                  // Node.Spec.ProviderID
};
// The reads (and the one write) the FTI clients issue against the API server.
// A missing object returns the apimachinery NotFound text, e.g.
// `metal3machines.infrastructure.cluster.x-k8s.io "machine-worker-0" not found`.
class ObjectStore {
public:
    virtual ~ObjectStore() {}
    virtual Error GetNode(const std::string& name, K8sObject* out) = 0;
    virtual Error GetMetal3Machine(const std::string& ns, const std::string& name, K8sObject* out) = 0;
    virtual Error GetBareMetalHost(const std::string& ns, const std::string& name, K8sObject* out) = 0;
    virtual Error ListNodeNames(std::vector<std::string>* out) = 0;
    // Status.DeviceID of every ComposableResource (cm/client.go:122-126)
    virtual Error ListComposableResourceDeviceIDs(std::vector<std::string>* out) = 0;
    // Status().Update after CM RemoveResource recorded a REMOVE_FAILED reason (cm/client.go:201-207)
    virtual Error UpdateStatus(const controller::ComposableResource& instance) = 0;
};
class TokenSource {   // fti.CachedToken.GetToken; the error text is surfaced verbatim ("unable to rotate token: ...")
public:
    virtual ~TokenSource() {}
    virtual Error GetToken() = 0;
};

// What the id_manager answered, turned into what CachedToken.GetToken returns (fti/token.go:72-175).  The POST
// itself — credentials from the Secret, TLS — stays with the host; the cache rule (a token is reused while
// expiry - 30 s is still ahead) and every error text are restated.
struct TokenReply {
    std::string secret_error;     // Secrets(...).Get failed: its text, e.g. `secrets "credentials" not found`
    std::string transport_error;  // client.PostForm failed
    int status = 200;
    std::string body;
};
// nil + *expiryUnix, or the error Token() returns (no "unable to rotate token: " prefix yet).
Error TokenFromReply(const TokenReply& r, long long* expiryUnix);
// base64.RawURLEncoding.DecodeString: false + Go's CorruptInputError text on bad input.
bool DecodeBase64RawURL(const std::string& in, std::string* out, std::string* err);

class ReplyTokenSource : public TokenSource {
public:
    ReplyTokenSource(const TokenReply& reply, long long nowUnix) : reply_(reply), now_(nowUnix) {}
    Error GetToken() override;
    int fetches = 0;              // how often the id_manager had to be asked

private:
    TokenReply reply_;
    long long now_;
    bool have_ = false;
    long long expiry_ = 0;
};

struct ClientConfig {
    std::string tenantID, clusterID;          // FTI_CDI_TENANT_ID, FTI_CDI_CLUSTER_ID
};

// Request bodies (json.Marshal of the reference's wire structs; also exported as cro_emit_*).
std::string FMScaleUpBody(const std::string& tenant, const std::string& machine, const std::string& type,
                          const std::string& model);                       // fm/api/scale_up.go:19-41
std::string FMScaleDownBody(const std::string& tenant, const std::string& machine, const std::string& type,
                            const std::string& resUUID);                   // fm/api/scale_down.go:19-41
std::string CMScaleUpBody(const std::string& specUUID, long long deviceCount);                              // cm/client.go:62-69
std::string CMScaleDownBody(const std::string& specUUID, long long deviceCount, const std::string& device); // cm/client.go:71-79

// ErrorBody decoding of a non-200 reply.  `what` is the operation word of the
// message ("scaleup", "scaledown", "get").
Error FMErrorFromReply(const std::string& what, const std::string& body);   // fm/client.go:172-181, 299-308, 495-503
Error CMErrorFromReply(const std::string& what, const std::string& body);   // cm/client.go:169-178, 249-257, 414-421
std::string formatFMErrorDetail(const std::string& code, const gojson::Value* rawMessage, const std::string& text);   // fm/client.go:60-72

// checkRemovingResources (cm/client.go:461-483) over the machine JSON.
struct CMRemovingResult {
    std::string specUUID;
    long long deviceCount = 0;
    Error err;
};
CMRemovingResult CMCheckRemovingResources(const std::string& machineBody, const std::string& specType,
                                          const std::string& specModel, const std::string& deviceID);

class FTIClientBase : public controller::CdiProvider {
public:
    FTIClientBase(const ClientConfig& cfg, Transport* t, ObjectStore* o, TokenSource* tok)
        : cfg_(cfg), http_(t), objects_(o), token_(tok) {}
    std::vector<HttpRequest> requests;        // every request issued, in order
    virtual Error GetResources(std::vector<DeviceInfo>* out) = 0;

protected:
    // metal3 walk shared by both flavours (cm/client.go:348-388, fm/client.go:415-450)
    Error machineIDFromAnnotations(const std::string& nodeName, bool useGivenNameInError, std::string* machineID);
    HttpReply send(const HttpRequest& r) { requests.push_back(r); return http_->Do(r); }
    ClientConfig cfg_;
    Transport* http_;
    ObjectStore* objects_;
    TokenSource* token_;
};

class FMClient : public FTIClientBase {       // internal/cdi/fti/fm/client.go
public:
    using FTIClientBase::FTIClientBase;
    Error AddResource(const controller::ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) override;
    Error RemoveResource(controller::ComposableResource& instance) override;
    Error CheckResource(const controller::ComposableResource& instance) override;
    Error GetResources(std::vector<DeviceInfo>* out) override;

private:
    Error getNodeMachineID(const std::string& nodeName, std::string* machineID);   // :415-464
    Error getMachineInfo(const std::string& machineID, std::string* body);         // :466-511
};

class CMClient : public FTIClientBase {       // internal/cdi/fti/cm/client.go
public:
    using FTIClientBase::FTIClientBase;
    Error AddResource(const controller::ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) override;
    Error RemoveResource(controller::ComposableResource& instance) override;
    Error CheckResource(const controller::ComposableResource& instance) override;
    Error GetResources(std::vector<DeviceInfo>* out) override;

private:
    Error getMachineInfo(const std::string& machineID, std::string* body);         // :390-430
};

// internal/cdi/sunfish/client.go:48-146 — one PATCH of a Redfish CompositionRequest; no ids come back
// (the reference's own TODO), CheckResource / GetResources are empty.  Plain HTTP, no token, no metal3 walk.
std::string SunfishBody(const std::string& name, long long count, const std::string& procType, const std::string& model);
class SunfishClient : public FTIClientBase {
public:
    explicit SunfishClient(Transport* t) : FTIClientBase(ClientConfig(), t, nullptr, nullptr) {}
    Error AddResource(const controller::ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) override;
    Error RemoveResource(controller::ComposableResource& instance) override;
    Error GetResources(std::vector<DeviceInfo>* out) override { out->clear(); return Error::Nil(); }

private:
    Error sendPatchRequest(const controller::ComposableResource& instance, long long count);   // :78-103
};

}  // namespace fabric
}  // namespace cro
