// token.go — cgo binding of cro_token_from_reply: what fti.CachedToken.Token makes of the id_manager's answer
// (internal/cdi/fti/token.go:138-175).  The POST itself stays where it is; only the reading of the reply moves.
//
// Source only: the Go toolchain is not in this repository's build image (INTEGRATION.md §1).
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -lcroprobe -ldl -lpthread
#include <stdlib.h>
#include "croprobe.h"
*/
import "C"

import (
	"encoding/json"
	"errors"
	"time"
	"unsafe"
)

// TokenReply is the id_manager's answer as the caller saw it.
type TokenReply struct {
	SecretError    string `json:"secret_error,omitempty"`
	TransportError string `json:"transport_error,omitempty"`
	Status         int    `json:"status"`
	Body           string `json:"body"`
}

// ExpiryFromReply returns the token's expiry, or the error CachedToken.Token would return
// (GetToken prefixes it with "unable to rotate token: ").
func ExpiryFromReply(r TokenReply) (time.Time, error) {
	in, err := json.Marshal(r)
	if err != nil {
		return time.Time{}, err
	}
	cin := C.CString(string(in))
	defer C.free(unsafe.Pointer(cin))
	const capacity = 16384
	buf := (*C.char)(C.malloc(capacity)) // caller-allocated; C does not retain it
	defer C.free(unsafe.Pointer(buf))
	var ln C.size_t
	if rc := C.cro_token_from_reply(cin, buf, capacity, &ln); rc != C.CRO_OK {
		return time.Time{}, errors.New(C.GoString(C.cro_strerror(rc)))
	}
	var out struct {
		Error  string `json:"error"`
		Expiry int64  `json:"expiry"`
	}
	if err := json.Unmarshal([]byte(C.GoStringN(buf, C.int(ln))), &out); err != nil {
		return time.Time{}, err
	}
	if out.Error != "" {
		return time.Time{}, errors.New(out.Error)
	}
	return time.Unix(out.Expiry, 0), nil
}
