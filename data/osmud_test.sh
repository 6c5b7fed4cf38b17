#!/bin/sh
# Kept at the reference's path (data/osmud_test.sh); the script itself lives in the package's tools/.
exec sh "$(dirname "$0")/../colearn_federated_learning_b200/tools/osmud_test.sh" "$@"
